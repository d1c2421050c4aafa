#!/bin/bash
# full verification of HEAD: GPU tests, bench line, ops bench, ncu capture of the exact step kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 600 gpurun_out/r2j_bench.err
python scripts/bench_ops.py > gpurun_out/r2j_ops.jsonl 2> gpurun_out/r2j_ops.err
cut -c1-220 gpurun_out/r2j_ops.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2j_x python scripts/profile_step.py --mode exact > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log
