"""Diagnostic: the data-parallel GraphedTrainStep under torchrun (2 ranks); dumps the Python stacks if it stalls.
    timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_graph_probe.py
"""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(45, exit=True)
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)


def log(*a):
    print("[rank %d %.1fs]" % (rank, time.time() - t0), *a, flush=True)


t0 = time.time()
from aesmc_b200 import train  # noqa: E402
from tests.models import nonlinear  # noqa: E402

torch.distributions.Distribution.set_default_validate_args(False)
torch.manual_seed(0)
init = nonlinear.Initial(dev)
trans, emis, prop = nonlinear.Transition(2.0).to(dev), nonlinear.Emission(0.03).to(dev), nonlinear.Proposal().to(dev)
params = list(train.get_chained_params(trans, emis, prop))
torch.manual_seed(200 + rank)
sampler = train.GraphedPriorSampler(init, nonlinear.Transition().to(dev), nonlinear.Emission().to(dev), 6, 8)
log("sampler captured")
opt = torch.optim.Adam(params, lr=1e-2, capturable=True)
# a bare all-reduce inside a graph first
x = torch.ones(4, device=dev)
dist.all_reduce(x)
torch.cuda.synchronize()
log("eager all-reduce ok", x.tolist())
mode = os.environ.get("CAPTURE_MODE", "thread_local")
g = torch.cuda.CUDAGraph()
y = torch.ones(4, device=dev)
with torch.cuda.graph(g, capture_error_mode=mode):
    dist.all_reduce(y)
log("bare all-reduce captured")
g.replay()
torch.cuda.synchronize()
log("bare all-reduce replayed", y.tolist())
step = train.GraphedTrainStep(sampler(clone=True), 64, "aesmc", init, trans, emis, prop, opt)
log("train step captured")
for i in range(3):
    loss = float(step(sampler()))
    log("replay", i, loss)
flat = torch.cat([p.detach().reshape(-1) for p in params])
ref = flat.clone()
dist.broadcast(ref, src=0)
log("in sync:", bool(torch.equal(flat, ref)))
step.release()
g.reset()
torch.cuda.synchronize()
log("graphs released")
dist.destroy_process_group()
log("process group destroyed")
faulthandler.cancel_dump_traceback_later()
