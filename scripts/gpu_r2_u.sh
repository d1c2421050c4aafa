#!/bin/bash
# what the driver does at round end: GPU tests, smoke(), bench (own arm and reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2u_pytest.log
tail -3 gpurun_out/r2u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
print("value %.4e ms %.2f frac %s e2e %.4e launches %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"]))
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ("error","fused_ms","eager_ms","graph_replay_ms_per_optimizer_step","ms_per_optimizer_step","index_mismatches")}) for k,v in d.items() if k in ("parity","infer_c3","train_c4","train_lgssm","clocks")})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
