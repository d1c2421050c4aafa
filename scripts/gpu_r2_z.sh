#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_step_parity_gpu.py tests/test_infer_gpu.py tests/test_ops_gpu.py tests/test_fused_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 400 python scripts/bench_sweep.py 2>/dev/null > gpurun_out/r2z_sweep.jsonl
python -c "
import sys,json
for l in open('gpurun_out/r2z_sweep.jsonl'):
    d=json.loads(l)
    print(d['B'],d['K'],d['mode'],'us/step %.1f'%d['us_per_step'],'frac %.3f'%d.get('frac_of_measured_hbm',0))"
