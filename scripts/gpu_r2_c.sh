#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_reference_unittests_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
echo "rc=$?" >> gpurun_out/r2c_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2c_x python scripts/profile_step.py --mode exact > gpurun_out/r2c_ncu.log 2>&1
AESMC_B200_LIB=$PWD/build/variants/libaesmc_alias5.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2c_x_alias5 python scripts/profile_step.py --mode exact >> gpurun_out/r2c_ncu.log 2>&1
tail -c 1500 gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_ncu.log
