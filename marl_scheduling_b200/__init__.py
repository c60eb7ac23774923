"""marl_scheduling_b200 -- B200-native batched rollout path of lr40/marl-scheduling.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute runs in
hand-written sm_100a CUDA kernels behind the C-ABI in include/msched.h.  No CPU fallback.
"""
from . import _lib
from ._lib import MschedError

__all__ = ["_lib", "MschedError", "BatchedSchedulingEnv"]


def __getattr__(name):
    if name == "BatchedSchedulingEnv":
        from .batched_env import BatchedSchedulingEnv
        return BatchedSchedulingEnv
    raise AttributeError(name)
