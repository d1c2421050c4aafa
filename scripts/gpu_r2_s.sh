#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -x -k "multi_cta or large" 2>&1 | tail -3
timeout 400 python scripts/bench_sweep.py 2>/dev/null > gpurun_out/r2s_sweep.jsonl
python -c "
import sys,json
for l in open('gpurun_out/r2s_sweep.jsonl'):
    d=json.loads(l)
    if d['K']>=100000: print(d['B'],d['K'],d['mode'],'us/step %.1f'%d['us_per_step'],'frac %.3f'%d.get('frac_of_measured_hbm',0))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:large_resample -s 1 -c 1 -f -o gpurun_out/r2s_resample python scripts/profile_step.py --mode fast --batch 64 --particles 1000000 --launches 3 > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log
