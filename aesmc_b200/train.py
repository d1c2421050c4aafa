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
    (distributions with validate_args=False).  Resampling uniforms come from torch's graph-safe generator on the
    device.

    Data-parallel (torch.distributed initialised with the nccl backend, one process per GPU, every rank constructing
    and calling the step in lock-step): the flattened gradient all-reduce of train() (aesmc/train.py:36 -> 37) is
    captured INSIDE the graph, between backward() and optimizer.step().  The eager warm-up steps create the NCCL
    communicator (communicator setup cannot be captured), and the capture runs in thread-local error mode: in the
    default global mode ProcessGroupNCCL's watchdog thread, which polls CUDA events of in-flight collectives from
    its own thread, invalidates the capture -- that, not NCCL itself, is what made the first attempt hang.
    ``global_batch_size``: total rows across ranks (default: local batch x world size)."""

    def __init__(self, observations, num_particles, algorithm, initial, transition, emission, proposal, optimizer,
                 resampling_mode=None, warmup_steps=3, global_batch_size=None):
        from . import inference
        if optimizer.defaults.get("capturable") is False:
            raise ValueError("construct the optimizer with capturable=True")
        self._world = distributed.world()[1]
        if self._world > 1 and torch.distributed.get_backend() != "nccl":
            raise NotImplementedError("a data-parallel GraphedTrainStep needs the nccl backend (gloo collectives cannot be captured)")
        self._params = [p for group in optimizer.param_groups for p in group["params"]]
        self._global_batch = global_batch_size
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
        if self._world > 1:
            torch.distributed.barrier()
        self.graph = torch.cuda.CUDAGraph()
        mode = {"capture_error_mode": "thread_local"} if self._world > 1 else {}
        with torch.cuda.graph(self.graph, **mode):
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
        if self._world > 1:
            distributed.all_reduce_gradients(self._params, self._B, self._global_batch or self._B * self._world)
        self._optimizer.step()
        return loss.detach()

    def release(self):
        """Free the captured graph.  Data-parallel: call this (on every rank) before
        torch.distributed.destroy_process_group() -- tearing the NCCL communicator down while an instantiated graph
        still holds its collective kernels blocks forever."""
        self.graph.reset()

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


class GraphedPriorSampler:
    """statistics.sample_from_prior (aesmc/statistics.py:108-162) for a CUDA model, captured ONCE as a CUDA graph and
    replayed per batch: the 2 T user-model calls and their torch.distributions sampling become one graph launch, with no
    host work and no host-device copy per batch -- the data half of SURVEY 8f-4, feeding GraphedTrainStep.

        sampler = GraphedPriorSampler(initial, transition, emission, num_timesteps=T, batch_size=B, device=dev)
        observations = sampler()        # list of T tensors [B, ...] on the device, fresh draws every call

    Every replay advances torch's graph-safe CUDA generator (torch.manual_seed controls the stream).  The returned
    tensors are the graph's static output buffers, overwritten by the next call (clone=True copies them).
    Requirements: callables free of host synchronisation (distributions with validate_args=False), tensor latents
    and observations (or dicts of tensors)."""

    def __init__(self, initial, transition, emission, num_timesteps, batch_size, device=None, keep_latents=False):
        from . import inference
        self._args = (initial, transition, emission, num_timesteps, batch_size)
        self._inference = inference
        probe = self._draw()[1][0]
        probe = next(iter(probe.values())) if isinstance(probe, dict) else probe
        if not probe.is_cuda:
            raise ValueError("GraphedPriorSampler needs a model on a CUDA device")
        self._dev = probe.device if device is None else torch.device(device)
        side = torch.cuda.Stream(device=self._dev)
        side.wait_stream(torch.cuda.current_stream(self._dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._draw()
        torch.cuda.current_stream(self._dev).wait_stream(side)
        torch.cuda.synchronize(self._dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            latents, observations = self._draw()
        squeeze = lambda v: v.squeeze(0)  # noqa: E731  (aesmc/train.py:63: a batch of one drops its batch axis)
        self.observations = [inference._map_tensors(squeeze, o) for o in observations]
        self.latents = latents if keep_latents else None

    def _draw(self):
        with torch.no_grad(), self._inference._scalars_by_fill_kernel():
            return statistics.sample_from_prior(*self._args)

    def __call__(self, clone=False):
        self.graph.replay()
        if not clone:
            return self.observations
        return [self._inference._map_tensors(lambda v: v.clone(), o) for o in self.observations]


class SyntheticDataset(torch.utils.data.Dataset):
    """Endless stream of observation sequences sampled from the generative model (aesmc/train.py:44-66).

    graphed=True (CUDA models): the ancestral sampling is captured once as a CUDA graph (GraphedPriorSampler) and every
    item is one graph replay -- the tensors of an item are then overwritten by the next item."""

    def __init__(self, initial, transition, emission, num_timesteps, batch_size, graphed=False):
        self.initial = initial
        self.transition = transition
        self.emission = emission
        self.num_timesteps = num_timesteps
        self.batch_size = batch_size
        self._sampler = GraphedPriorSampler(initial, transition, emission, num_timesteps, batch_size) if graphed else None

    def __getitem__(self, index):
        if self._sampler is not None:
            return self._sampler()
        _, observations = statistics.sample_from_prior(self.initial, self.transition, self.emission,
                                                       self.num_timesteps, self.batch_size)
        return [o.detach().squeeze(0) for o in observations]

    def __len__(self):
        return sys.maxsize


def get_synthetic_dataloader(initial, transition, emission, num_timesteps, batch_size, graphed=False):
    return torch.utils.data.DataLoader(
        SyntheticDataset(initial, transition, emission, num_timesteps, batch_size, graphed=graphed),
        batch_size=1, collate_fn=lambda items: items[0])
