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
