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


def pin_host_to_gpu(gpu_index):
    """Best effort: bind this process (and therefore the pinned host buffers it allocates afterwards, first touch) to
    the CPUs of the NUMA node the GPU hangs off (NVML's ideal CPU affinity).  The host-buffer step writes results
    from the SMs straight into host memory; with all ranks' buffers on one node the far GPUs cross the socket link.
    Returns a short description (what was set, or why nothing was)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            node = None
        if not cpus:
            return f"numa node {node}: no usable CPUs in the GPU's affinity mask"
        if cpus == allowed:
            return f"numa node {node}: the GPU's affinity mask is every allowed CPU ({len(allowed)}): nothing to pin"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node}: bound to {len(cpus)} of {len(allowed)} CPUs"
    except Exception as e:  # no NVML, no permission: run unpinned
        return f"unpinned ({type(e).__name__})"


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
