#!/bin/bash
# N-GPU bench (all legs) under torchrun, tight timeout
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2y_bench_${N}gpu.json 2> gpurun_out/r2y_bench_${N}gpu.err
echo "rc=$?"
tail -c 500 gpurun_out/r2y_bench_${N}gpu.err
python - <<PY
import json
txt=open("gpurun_out/r2y_bench_${N}gpu.json").read().strip().splitlines()
d=json.loads(txt[-1])
print(d["n_gpus"], "value %.3e"%d["value"], "ms %.2f"%d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
for k in ("strong","sweep_c5","infer_c3","train_c4","train_lgssm"):
    v=d.get(k)
    if isinstance(v,list): 
        for e in v: print("  ",k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in e.items() if a in("K","mode","rows_total","us_per_step","value","frac_of_measured_hbm")})
    else: print("  ",k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in("model","graph_replay_includes","note","unit","unit_ms")})
PY
