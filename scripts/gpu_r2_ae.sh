#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_r2_k.sh r2ae ncu 2>&1 | tail -4
python scripts/bench_step_variant.py --mode fast --label fast 2>/dev/null | cut -c1-200
