"""Batched PPO rollout pieces over the C-ABI: actor forward (+ categorical sample / log-prob)
and discounted returns.  Mirrors reference src/PPOmodules.py:25-72 (ActorCritic.act),
:114-125 (PPO.selectAction) and :128-137 (returns prologue of PPO.update)."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L


def param_count(n_in, n_hidden, n_actions):
    return L.lib().msched_mlp_param_count(n_in, n_hidden, n_actions)


def pack_actor(state_dict, prefix="actor."):
    """Flatten an ActorCritic.actor state dict (Linear at indices 0, 2, 4) into the ABI's
    per-net layout [W1 | b1 | W2 | b2 | W3 | b3] (float32, torch [out][in] order)."""
    parts = []
    for i in (0, 2, 4):
        parts.append(torch.as_tensor(state_dict[f"{prefix}{i}.weight"]).float().reshape(-1))
        parts.append(torch.as_tensor(state_dict[f"{prefix}{i}.bias"]).float().reshape(-1))
    return torch.cat(parts)


class MlpGroup:
    """n_nets identically shaped actor MLPs with their weights resident on the device."""

    def __init__(self, n_in, n_hidden, n_actions, weights, device, unit_div=1):
        self.n_in, self.n_hidden, self.n_actions = n_in, n_hidden, n_actions
        w = torch.as_tensor(weights, dtype=torch.float32)
        pc = param_count(n_in, n_hidden, n_actions)
        w = w.reshape(-1, pc)
        self.n_nets = w.shape[0]
        self.weights = w.to(device).contiguous()
        self.unit_div = unit_div
        self.desc = L.MschedMlpGroup(n_in, n_hidden, n_actions, self.n_nets, unit_div, 0,
                                     self.weights.data_ptr())

    @classmethod
    def from_state_dicts(cls, n_in, n_hidden, n_actions, state_dicts, device, prefix="actor."):
        return cls(n_in, n_hidden, n_actions,
                   torch.stack([pack_actor(sd, prefix) for sd in state_dicts]), device)

    @classmethod
    def random(cls, n_in, n_hidden, n_actions, n_nets, device, seed=0):
        """torch.nn.Linear default initialisation (what the reference's ActorCritic gets)."""
        g = torch.Generator().manual_seed(seed)
        nets = []
        for _ in range(n_nets):
            parts = []
            for fan_in, fan_out in ((n_in, n_hidden), (n_hidden, n_hidden), (n_hidden, n_actions)):
                bound = 1.0 / (fan_in ** 0.5)
                parts.append((torch.rand(fan_out * fan_in, generator=g) * 2 - 1) * bound)
                parts.append((torch.rand(fan_out, generator=g) * 2 - 1) * bound)
            nets.append(torch.cat(parts))
        return cls(n_in, n_hidden, n_actions, torch.stack(nets), device)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _make_io(x, x_stride, units, n_envs, env_stride, seed, step, row_offset, u, action, logprob, probs,
             action_rec, action_rec_stride, gather_core, n_cores, x_used, timeline, step_dev):
    io = L.MschedActorIO()
    io.x, io.x_stride, io.units, io.n_envs, io.n_cores = x.data_ptr(), x_stride, units, n_envs, n_cores
    io.env_stride, io.row_offset, io.seed, io.step = env_stride, row_offset, seed, step
    io.u_override = None if u is None else u.data_ptr()
    io.action = None if action is None else action.data_ptr()
    io.logprob = None if logprob is None else logprob.data_ptr()
    io.probs = None if probs is None else probs.data_ptr()
    io.action_rec = None if action_rec is None else action_rec.data_ptr()
    io.action_rec_stride = action_rec_stride
    io.gather_core = None if gather_core is None else gather_core.data_ptr()
    io.x_used = None if x_used is None else x_used.data_ptr()
    io.timeline = None if timeline is None else timeline.data_ptr()
    io.step_dev = None if step_dev is None else step_dev.data_ptr()  # int64 device counter (graph replays)
    return io


def actor_forward(group, x, x_stride, units, n_envs, env_stride=0, seed=0, step=0, row_offset=0,
                  u=None, want_probs=False, action=None, logprob=None, action_rec=None,
                  action_rec_stride=0, gather_core=None, n_cores=0, x_used=None, timeline=None, step_dev=None):
    """x: int16 device tensor; unit u of env b is the row at x[b*env_stride + u*x_stride : +n_in].
    action_rec (int16 view into the env's action record) additionally receives the action the
    world is handed; gather_core (int32 [n_envs*units]) turns the launch into the free-price price
    chooser (its 4 inputs are sliced out of the offer observation row, src/PPOmodules.py:312-332);
    x_used (int16 [n_envs*units, n_in]) receives the inputs actually fed (PPO buffer.states).
    Returns (action int32 [n_envs*units], logprob float32, probs or None)."""
    dev = x.device
    M = n_envs * units
    if action is None:
        action = torch.empty(M, dtype=torch.int32, device=dev)
    if logprob is None:
        logprob = torch.empty(M, dtype=torch.float32, device=dev)
    probs = torch.empty((M, group.n_actions), dtype=torch.float32, device=dev) if want_probs else None
    if u is not None:
        u = torch.as_tensor(u, dtype=torch.float32).to(dev).contiguous()
    io = _make_io(x, x_stride, units, n_envs, env_stride, seed, step, row_offset, u, action, logprob, probs,
                  action_rec, action_rec_stride, gather_core, n_cores, x_used, timeline, step_dev)
    L.check(L.lib().msched_actor_forward(C.byref(group.desc), C.byref(io), _stream(dev)))
    return action, logprob, probs


def returns(rewards, gamma, normalise=True, out=None):
    """rewards float32 [T][M] on the device -> (normalised) Monte-Carlo returns, same shape."""
    r = rewards.contiguous()
    T, M = r.shape
    if out is None:
        out = torch.empty_like(r)
    L.check(L.lib().msched_returns(r.data_ptr(), T, M, float(gamma), int(normalise), out.data_ptr(),
                                   _stream(r.device)))
    return out


def ppo_grad_supported(n_in, n_hidden, n_actions):
    """Shapes msched_ppo_grad serves (the divided / shared 16-wide nets); others use autograd."""
    return n_hidden == 16 and 1 <= n_in <= 64 and 1 <= n_actions <= 16


def ppo_grad(actor_w, critic_w, n_in, n_actions, x, action, logprob_old, returns, net_ids, unit_ids,
             grad_actor, grad_critic, eps_clip=0.2, entropy_coef=0.01, value_coef=0.5, stats=None, workspace=None):
    """One epoch's PPO gradient (src/PPOmodules.py:139-174) for the selected nets, forward + backward in
    one kernel.  actor_w / critic_w: float32 [n_nets, pc] on the device; x: int16 [TB, U, n_in] view (any
    tb / unit strides, innermost contiguous); action int32 / logprob_old / returns float32: contiguous
    [TB, U]; net_ids int32 [n_sel], unit_ids int32 [n_sel, m] on the device.  Rows net_ids of grad_actor /
    grad_critic are overwritten.  Returns (stats [n_sel, 4], workspace) for reuse."""
    dev = x.device
    TB, U = x.shape[0], x.shape[1]
    assert x.dtype == torch.int16 and x.stride(2) == 1 and x.shape[2] == n_in
    for t in (action, logprob_old, returns):
        assert t.is_contiguous() and t.numel() == TB * U
    assert action.dtype == torch.int32 and net_ids.dtype == torch.int32 and unit_ids.dtype == torch.int32
    b = L.MschedPpoBatch()
    b.actor_weights, b.critic_weights = actor_w.data_ptr(), critic_w.data_ptr()
    b.n_in, b.n_hidden, b.n_actions, b.n_nets = n_in, 16, n_actions, actor_w.shape[0]
    b.x, b.x_tb_stride, b.x_unit_stride = x.data_ptr(), x.stride(0), x.stride(1)
    b.action, b.logprob_old, b.returns = action.data_ptr(), logprob_old.data_ptr(), returns.data_ptr()
    b.n_tb, b.units, b.n_sel, b.units_per_net = TB, U, unit_ids.shape[0], unit_ids.shape[1]
    b.net_ids, b.unit_ids = net_ids.data_ptr(), unit_ids.data_ptr()
    b.eps_clip, b.entropy_coef, b.value_coef = eps_clip, entropy_coef, value_coef
    b.grad_actor, b.grad_critic = grad_actor.data_ptr(), grad_critic.data_ptr()
    if stats is None:
        stats = torch.empty((unit_ids.shape[0], 4), dtype=torch.float32, device=dev)
    b.stats = stats.data_ptr()
    need = C.c_uint64()
    L.check(L.lib().msched_ppo_workspace_bytes(C.byref(b), C.byref(need)))
    if workspace is None or workspace.numel() * 4 < need.value:
        workspace = torch.empty((need.value + 3) // 4, dtype=torch.float32, device=dev)
    b.workspace, b.workspace_bytes = workspace.data_ptr(), workspace.numel() * 4
    L.check(L.lib().msched_ppo_grad(C.byref(b), _stream(dev)))
    return stats, workspace


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam's step over a flat float32 device buffer, in place (step counts from 1)."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == param.numel()
    L.check(L.lib().msched_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                     param.numel(), float(lr), beta1, beta2, eps, int(step), _stream(param.device)))


def policy_step_group(group, units, x_offset, x_stride, rec_offset, seed, action=None, logprob=None, x_used=None,
                      u=None, probs=None):
    """One MschedPolicyGroup of msched_policy_step: `group` (MlpGroup) serves `units` rows per environment whose
    observation rows start at x_offset + u * x_stride (int16 elements inside the observation record) and whose
    actions go to rec_offset + u of the action record.  action int32 / logprob float32 [B, units]; x_used int16
    [B, units, stride] (the row as read: word aligned, one leading pad when x_offset is odd); u float32 [B, units]
    draw overrides; probs float32 [B, units, A]."""
    g = L.MschedPolicyGroup()
    g.nets = group.desc
    g.units, g.x_offset, g.x_stride, g.rec_offset, g.seed = units, x_offset, x_stride, rec_offset, seed
    g.action = None if action is None else action.data_ptr()
    g.logprob = None if logprob is None else logprob.data_ptr()
    if x_used is not None:
        assert x_used.dtype == torch.int16 and x_used.stride(-1) == 1
        g.x_used, g.x_used_stride = x_used.data_ptr(), x_used.stride(-2)
    g.u_override = None if u is None else u.data_ptr()
    g.probs = None if probs is None else probs.data_ptr()
    g._keep = (group, action, logprob, x_used, u, probs)
    return g


def policy_step_supported(acceptor, core, price=None):
    """Net shapes msched_policy_step has kernels for (the BASELINE domains' 16-wide nets)."""
    if acceptor.n_hidden != 16 or core.n_hidden != 16 or (price is not None and price.n_hidden != 16):
        return False
    shapes = {(15, 8, True), (27, 10, False), (15, 8, False), (11, 8, False)}
    return ((acceptor.n_in, core.n_in, price is not None) in shapes and acceptor.n_actions <= (16 if acceptor.n_in == 27 else 8)
            and core.n_actions <= 8 and (price is None or (price.n_in == 4 and price.n_actions <= 16)))


def policy_step(obs, obs_stride, n_envs, n_cores, acceptor, core, price=None, action_rec=None, action_rec_stride=0,
                env_offset=0, step=0, step_dev=None, input_bound=0):
    """Every PPO unit of a rollout step in ONE launch (msched_policy_step): acceptor / core / price are
    MschedPolicyGroup objects from policy_step_group (price None = fixed prices)."""
    ps = L.MschedPolicyStep()
    ps.obs, ps.obs_stride, ps.n_envs, ps.n_cores = obs.data_ptr(), obs_stride, n_envs, n_cores
    ps.action_rec = None if action_rec is None else action_rec.data_ptr()
    ps.action_rec_stride, ps.env_offset, ps.step = action_rec_stride, env_offset, step
    ps.step_dev = None if step_dev is None else step_dev.data_ptr()
    ps.input_bound = int(input_bound)  # max |observation value| (0 = unknown): 1..511 allows the tensor-core kernel
    ps.acceptor, ps.core = acceptor, core
    if price is not None:
        ps.price = price
    L.check(L.lib().msched_policy_step(C.byref(ps), _stream(obs.device)))
