"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Environments shard trivially (no cross-env state anywhere in the reference: one World per
process, src/world.py:210-254), so the env step needs NO collective: rank g owns the contiguous
global env range [g*B/G, (g+1)*B/G) and its Philox counters use the GLOBAL env index, so results do
not depend on the number of GPUs.  The only exchange is the PPO gradient all-reduce when nets are
shared across environments (config 4, GloballySharedPPO.update src/PPOmodules.py:395-449 is the
call-site analogue): all gradients are packed into one flat buffer and reduced with ONE
all_reduce per epoch -- KB-sized, latency-bound on NVLink5/NVSwitch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_envs(total_envs, world_size, rank):
    """Contiguous shard of the global env range: returns (env_offset, n_envs)."""
    base, rem = divmod(total_envs, world_size)
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_gradients(params):
    """Average the gradients of `params` over all ranks with a single flat-buffer all_reduce."""
    if not is_distributed():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n


def allreduce_mean_(flat):
    """In-place average of one flat gradient buffer over all ranks (a single all_reduce)."""
    if not is_distributed():
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    return flat


def max_over_ranks(value, device="cpu"):
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
