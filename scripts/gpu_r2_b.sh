#!/bin/bash
# round 2, GPU call B: new exact row kernel -- parity tests, then timing old vs new
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_reference_unittests_gpu.py tests/test_infer_gpu.py -m gpu -q -x 2>&1 | tail -60 > gpurun_out/r2b_pytest.log
echo "rc=$?" >> gpurun_out/r2b_pytest.log
: > gpurun_out/r2b_variants.jsonl
AESMC_DISABLE_STEP_X=1 python scripts/bench_step_variant.py --label old_reg_kernel >> gpurun_out/r2b_variants.jsonl 2>gpurun_out/r2b_err.log
python scripts/bench_step_variant.py --label x_kernel >> gpurun_out/r2b_variants.jsonl 2>>gpurun_out/r2b_err.log
for f in build/variants/*.so; do
  [ -f "$f" ] && AESMC_B200_LIB=$PWD/$f python scripts/bench_step_variant.py --label $(basename $f .so) >> gpurun_out/r2b_variants.jsonl 2>>gpurun_out/r2b_err.log
done
tail -c 2500 gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_variants.jsonl
tail -5 gpurun_out/r2b_err.log
