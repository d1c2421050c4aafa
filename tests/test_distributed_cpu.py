"""Data-parallel plumbing on CPU: world_size-2 gloo process group (no GPU).  Covers batch-row sharding
by global index, the flattened gradient all-reduce used by train(), and the evidence reductions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aesmc_b200 import distributed


def test_shard_bounds_cover_batch_without_overlap():
    for B in (1, 2, 7, 8, 4096, 4099):
        for W in (1, 2, 3, 8):
            spans = [distributed.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_helpers_single_process():
    obs = [torch.arange(10.0) + t for t in range(3)]
    part = distributed.shard_batch(obs, rank=1, world_size=3)
    assert [p.tolist() for p in part] == [[4 + t, 5 + t, 6 + t] for t in range(3)]
    d = distributed.shard_batch({"y": torch.arange(10.0)}, rank=0, world_size=3)
    assert d["y"].tolist() == [0, 1, 2, 3]
    u = np.arange(20.0).reshape(2, 10)
    assert distributed.shard_uniforms(u, rank=2, world_size=3).tolist() == [[7, 8, 9], [17, 18, 19]]
    assert distributed.world() == (0, 1) and not distributed.is_active()
    x = torch.arange(4.0)
    assert distributed.gather_rows(x, 4) is x
    assert float(distributed.global_mean(x, 4)) == 1.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Linear(3, 1)
        full_x = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3) / 10
        full_y = torch.arange(B, dtype=torch.float32)
        x = distributed.shard_batch(full_x)
        y = distributed.shard_batch(full_y)
        loss = ((model(x).squeeze(-1) - y) ** 2).mean()      # local batch-mean, like losses.get_loss
        loss.backward()
        distributed.all_reduce_gradients(list(model.parameters()), x.shape[0], B)
        per_row = (model(x).squeeze(-1) - y).detach()
        gathered = distributed.gather_rows(per_row, B)
        mean = distributed.global_mean(per_row, B)
        # the captured training step refuses to run data-parallel on gloo (only NCCL collectives can be captured)
        from aesmc_b200 import train
        try:
            train.GraphedTrainStep([x], 4, "aesmc", None, None, None, None, torch.optim.SGD(model.parameters(), lr=0.1))
            refused = False
        except NotImplementedError:
            refused = True
        if rank == 0:
            torch.save({"grad_w": model.weight.grad, "grad_b": model.bias.grad, "rows": gathered, "mean": mean,
                        "refused": refused}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_gradients_equal_single_process(tmp_path, B):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Linear(3, 1)
    x = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3) / 10
    y = torch.arange(B, dtype=torch.float32)
    ((model(x).squeeze(-1) - y) ** 2).mean().backward()
    torch.testing.assert_close(got["grad_w"], model.weight.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["grad_b"], model.bias.grad, rtol=1e-5, atol=1e-6)
    rows = (model(x).squeeze(-1) - y).detach()
    torch.testing.assert_close(got["rows"], rows)            # global row order, ragged shards included
    torch.testing.assert_close(got["mean"], rows.mean())
    assert got["refused"]
