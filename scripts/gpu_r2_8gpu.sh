#!/bin/bash
# eight B200s: bench.py under torchrun, tight timeout
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 > gpurun_out/r2g_bench_8gpu.json 2> gpurun_out/r2g_bench_8gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_8gpu.json").read().strip().splitlines()[-1])
    print("N=8 value %.4e ms %.2f frac %s e2e %.4e strong %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],{k:d["strong"].get(k) for k in ("value","ms_per_step")}))
    print({k:d["train_c4"].get(k) for k in ("ms_per_optimizer_step","graph_replay_ms_per_optimizer_step","replicas_in_sync","graph_replicas_in_sync")})
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/r2g_bench_8gpu.err").read()[-1500:])
PY
