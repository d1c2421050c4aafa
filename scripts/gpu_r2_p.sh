#!/bin/bash
# multi-CTA path: parity tests, config-5 sweep, launch list at B = 64, K = 1e6
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_parity_gpu.py -m gpu -q -x -k "multi_cta or large" 2>&1 | tail -12 > gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log
timeout 400 python scripts/bench_sweep.py > gpurun_out/r2p_sweep.jsonl 2> gpurun_out/r2p_sweep.err
python - <<PY
import json
for l in open("gpurun_out/r2p_sweep.jsonl"):
    d=json.loads(l)
    if d["K"]>=100000: print(d["B"],d["K"],d["mode"],"us/step %.1f"%d["us_per_step"], "frac %.3f"%d.get("frac_of_measured_hbm",0))
PY
tail -2 gpurun_out/r2p_sweep.err
