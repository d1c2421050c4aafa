"""Minimal optimisation loop and the synthetic data source (mirrors aesmc/train.py:10-71).

Under torch.distributed (one process per GPU) each rank trains on its own shard of batch rows and
the parameter gradients are all-reduced between backward() and step(); callbacks fire on rank 0 only.
"""
import itertools
import sys

import torch
import torch.nn as nn
import torch.utils.data

from . import distributed
from . import losses
from . import statistics


def get_chained_params(*objects):
    """Iterator over the parameters of every nn.Module among ``objects`` (None if there is none)."""
    modules = [o for o in objects if o is not None and isinstance(o, nn.Module)]
    if not modules:
        return None
    return itertools.chain.from_iterable(m.parameters() for m in modules)


def train(dataloader, num_particles, algorithm, initial, transition, emission, proposal, num_epochs,
          num_iterations_per_epoch=None, optimizer_algorithm=torch.optim.Adam, optimizer_kwargs={},
          callback=None, global_batch_size=None):
    """Run num_epochs passes over ``dataloader``; each batch is one optimiser step on get_loss().

    global_batch_size: total rows across ranks when running data-parallel with per-rank shards of
    unequal size (default: local batch x world size)."""
    parameters = list(get_chained_params(initial, transition, emission, proposal) or [])
    optimizer = optimizer_algorithm(parameters, **optimizer_kwargs)
    rank, world_size = distributed.world()
    for epoch_idx in range(num_epochs):
        for epoch_iteration_idx, observations in enumerate(dataloader):
            if num_iterations_per_epoch is not None and epoch_iteration_idx == num_iterations_per_epoch:
                break
            optimizer.zero_grad()
            loss = losses.get_loss(observations, num_particles, algorithm, initial, transition, emission,
                                   proposal)
            loss.backward()
            if world_size > 1:
                first = observations[0]
                local = (next(iter(first.values())) if isinstance(first, dict) else first).size(0)
                distributed.all_reduce_gradients(parameters, local, global_batch_size or local * world_size)
            optimizer.step()
            if callback is not None and rank == 0:
                callback(epoch_idx, epoch_iteration_idx, loss, initial, transition, emission, proposal)


class GraphedTrainStep:
    """One optimisation step -- get_loss() forward, backward, optimizer.step() -- captured once as a CUDA graph
    and replayed per batch: at small and medium shapes a training step is bound by the hundreds of launches
    torch issues for the user model and its autograd graph, not by the GPU.

        opt = torch.optim.Adam(params, lr=1e-3, capturable=True)
        step = GraphedTrainStep(observations, num_particles, 'aesmc', initial, transition, emission, proposal, opt)
        for batch in data:
            loss = step(batch)            # 0-d CUDA tensor, overwritten by the next call

    The constructor runs three ordinary (eager) steps on ``observations`` first -- they do train the model --
    so that the optimizer state and autograd's lazily created buffers exist before the capture, as torch's
    whole-network capture recipe requires.  Requirements: CUDA model and observations of fixed shapes, a
    capturable optimizer (Adam-family: capturable=True), callables free of host synchronisation
    (distributions with validate_args=False), a single process (no gradient all-reduce inside the graph).
    Resampling uniforms come from torch's graph-safe generator on the device."""

    def __init__(self, observations, num_particles, algorithm, initial, transition, emission, proposal, optimizer,
                 resampling_mode=None, warmup_steps=3):
        from . import inference
        if distributed.world()[1] > 1:
            raise NotImplementedError("GraphedTrainStep does not capture the gradient all-reduce; use train() under torch.distributed")
        if optimizer.defaults.get("capturable") is False:
            raise ValueError("construct the optimizer with capturable=True")
        self._args = (num_particles, algorithm, initial, transition, emission, proposal)
        self._mode = resampling_mode
        self._optimizer = optimizer
        self._inference = inference
        self.observations = [inference._map_tensors(lambda v: v.detach().clone(), o) for o in observations]
        probe = inference._first_tensor(self.observations[0])
        if not probe.is_cuda:
            raise ValueError("GraphedTrainStep needs CUDA observations")
        self._dev, self._T, self._B = probe.device, len(self.observations), probe.size(0)
        side = torch.cuda.Stream(device=self._dev)
        side.wait_stream(torch.cuda.current_stream(self._dev))
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                self._step()
        torch.cuda.current_stream(self._dev).wait_stream(side)
        torch.cuda.synchronize(self._dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _step(self):
        num_particles, algorithm, initial, transition, emission, proposal = self._args
        kwargs = {"check_finite": False, "_allow_fused": False, "resampling_mode": self._mode}
        if algorithm == "aesmc" and self._T > 1:
            kwargs["uniforms"] = torch.rand(self._T - 1, self._B, dtype=torch.float64, device=self._dev)
        self._optimizer.zero_grad(set_to_none=True)
        with self._inference._scalars_by_fill_kernel():
            loss = losses.get_loss(self.observations, num_particles, algorithm, initial, transition, emission, proposal,
                                   **kwargs)
            loss.backward()
        self._optimizer.step()
        return loss.detach()

    def __call__(self, observations=None):
        if observations is not None:
            if len(observations) != self._T:
                raise ValueError("expected %d observations, got %d" % (self._T, len(observations)))
            for dst, src in zip(self.observations, observations):
                if isinstance(dst, dict):
                    for name in dst:
                        dst[name].copy_(src[name], non_blocking=True)
                else:
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


class SyntheticDataset(torch.utils.data.Dataset):
    """Endless stream of observation sequences sampled from the generative model."""

    def __init__(self, initial, transition, emission, num_timesteps, batch_size):
        self.initial = initial
        self.transition = transition
        self.emission = emission
        self.num_timesteps = num_timesteps
        self.batch_size = batch_size

    def __getitem__(self, index):
        _, observations = statistics.sample_from_prior(self.initial, self.transition, self.emission,
                                                       self.num_timesteps, self.batch_size)
        return [o.detach().squeeze(0) for o in observations]

    def __len__(self):
        return sys.maxsize


def get_synthetic_dataloader(initial, transition, emission, num_timesteps, batch_size):
    return torch.utils.data.DataLoader(
        SyntheticDataset(initial, transition, emission, num_timesteps, batch_size),
        batch_size=1, collate_fn=lambda items: items[0])
