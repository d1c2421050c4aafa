#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -x -k "multi_cta or large" 2>&1 | tail -3
for lib in "" build/variants/libaesmc_rs6.so build/variants/libaesmc_rs8.so; do
  echo "== lib: ${lib:-default}"
  AESMC_B200_LIB=${lib:+$PWD/$lib} timeout 400 python scripts/bench_sweep.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['K']>=100000: print(d['B'],d['K'],d['mode'],'us/step %.1f'%d['us_per_step'],'frac %.3f'%d.get('frac_of_measured_hbm',0))"
done
