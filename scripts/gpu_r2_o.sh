#!/bin/bash
# bench at N = 1 and (with 2 GPUs) N = 2, tight timeouts
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench_1gpu.json 2> gpurun_out/r2o_bench_1gpu.err
tail -c 400 gpurun_out/r2o_bench_1gpu.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err
tail -c 600 gpurun_out/r2o_bench_2gpu.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_bench_*gpu.json")):
    txt=open(f).read().strip().splitlines()
    if not txt: print(f,"EMPTY"); continue
    d=json.loads(txt[-1])
    print(f, d["n_gpus"], "value %.3e"%d["value"], "ms %.2f"%d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
    for k in ("infer_c3","train_c4","train_lgssm","strong"):
        v=d.get(k,{})
        print("  ",k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in("model","graph_replay_includes","note")})
PY
