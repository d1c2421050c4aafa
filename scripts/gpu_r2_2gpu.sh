#!/bin/bash
# two B200s: the NCCL test and bench.py under torchrun (tight timeouts: a hung rendezvous must not eat the budget)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2g_test_multi_gpu.log; cat gpurun_out/r2g_test_multi_gpu.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2 value %.4e ms %.2f frac %s e2e %.4e strong %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],{k:d["strong"].get(k) for k in ("value","ms_per_step")}))
    print({k:d["train_c4"].get(k) for k in ("ms_per_optimizer_step","graph_replay_ms_per_optimizer_step","replicas_in_sync","graph_replicas_in_sync")})
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/r2g_bench_2gpu.err").read()[-1500:])
PY
