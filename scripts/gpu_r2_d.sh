#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2e_pytest.log
echo "rc=$?" >> gpurun_out/r2e_pytest.log
: > gpurun_out/r2e_variants.jsonl
python scripts/bench_step_variant.py --label x_kernel >> gpurun_out/r2e_variants.jsonl 2>gpurun_out/r2e_err.log
for f in build/variants/*.so; do
  [ -f "$f" ] && AESMC_B200_LIB=$PWD/$f python scripts/bench_step_variant.py --label $(basename $f .so) >> gpurun_out/r2e_variants.jsonl 2>>gpurun_out/r2e_err.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2e_x python scripts/profile_step.py --mode exact > gpurun_out/r2e_ncu.log 2>&1
tail -c 800 gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_variants.jsonl
tail -3 gpurun_out/r2e_err.log
