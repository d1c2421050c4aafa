#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_infer_gpu.py -m gpu -q -x 2>&1 | tail -3
for lib in "" build/variants/libaesmc_fast6.so build/variants/libaesmc_fast0.so; do
  AESMC_B200_LIB=${lib:+$PWD/$lib} python scripts/bench_step_variant.py --mode fast --label "${lib:-default}" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['label'], 'fast us(s=1) %.2f us(s=8) %.2f  frac %.4f mism %d/%d gather %s'%(d['us_scale_1'],d['us_scale_8'],d['frac_of_6550_scale_1'],d['mismatch_scale_1'],d['mismatch_scale_8'],d['gather_ok_scale_1']))"
done
