#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","roofline","e2e","e2e_eager","e2e_eager_graph","parity","clocks"):
    v=d.get(k)
    if isinstance(v,dict): v={a:b for a,b in v.items() if a not in("api","against","note","model","tolerance")}
    print(k,v)
PY
tail -3 gpurun_out/r2h_bench.err
