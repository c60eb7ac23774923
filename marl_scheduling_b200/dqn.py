"""Batched DQN units: the interface of reference src/DQNmodules.py
(`ReplayMemory.push/sample`, `DQNEntity.selectAction`, `optimize_model`) and of
`DividedFixPriceDQNAgent` (src/Agent.py:303-356) with a leading environment dimension.

Rollout: the epsilon-greedy Q-forward of every (environment, unit) runs in the CUDA kernel behind
msched_dqn_select, straight on the observation record, writing the action record.  Learning
(`optimize_model`: SmoothL1Loss against the target net, gradients clamped to [-1, 1], Adam) runs in
msched_dqn_grad (forward of both nets + backward of every unit's Q-net in one launch, bit-reproducible) followed
by msched_adam_step, over device-resident replay memories.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ReplayMemory:
    """Device ring buffer of transitions (state, action, next_state, reward) per unit
    (src/DQNmodules.py:13-31).  `push` takes one transition per environment and unit."""

    def __init__(self, capacity, units, n_in, device):
        self.capacity, self.units, self.n_in = int(capacity), units, n_in
        self.state = torch.zeros((self.capacity, units, n_in), dtype=torch.int16, device=device)
        self.next_state = torch.zeros_like(self.state)
        self.action = torch.zeros((self.capacity, units), dtype=torch.int64, device=device)
        self.reward = torch.zeros((self.capacity, units), dtype=torch.float32, device=device)
        self.size = 0
        self.next = 0

    def push(self, state, action, next_state, reward):
        """state / next_state int16 [B, units, n_in]; action [B, units]; reward [B, units]."""
        B = state.shape[0]
        if B > self.capacity:
            # more transitions than the ring holds: only the last `capacity` survive a sequential push, and
            # keeping exactly those leaves no duplicate indices (index_put with duplicates is unordered)
            self.next = (self.next + B - self.capacity) % self.capacity
            state, next_state = state[-self.capacity:], next_state[-self.capacity:]
            action, reward = action[-self.capacity:], reward[-self.capacity:]
            B = self.capacity
        # at most two wrap-free contiguous slice copies
        first = min(B, self.capacity - self.next)
        for dst, src in ((self.state, state), (self.next_state, next_state), (self.action, action.long()),
                         (self.reward, reward.float())):
            dst[self.next: self.next + first] = src[:first]
            if first < B:
                dst[: B - first] = src[first:]
        self.next = (self.next + B) % self.capacity
        self.size = min(self.capacity, self.size + B)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.state.device, generator=generator)
        return self.state[idx], self.action[idx], self.next_state[idx], self.reward[idx]

    def __len__(self):
        return self.size


class BatchedDQN:
    """n_nets Q-nets Linear(in,16)-Tanh-Linear(16,A) (DQNEntity), policy + target, one per unit."""
    H = 16

    def __init__(self, n_in, n_actions, units, run_start, run_end, run_decay, gamma, memory_size, device, seed=0):
        self.n_in, self.A, self.units = n_in, n_actions, units
        self.RUN_START, self.RUN_END, self.RUN_DECAY = run_start, run_end, run_decay
        self.gamma = gamma
        self.device = device
        g = torch.Generator().manual_seed(seed)
        nets = []
        for _ in range(units):
            parts = []
            for fan_in, fan_out in ((n_in, self.H), (self.H, n_actions)):
                bound = 1.0 / (fan_in ** 0.5)
                parts.append((torch.rand(fan_out * fan_in, generator=g) * 2 - 1) * bound)
                parts.append((torch.rand(fan_out, generator=g) * 2 - 1) * bound)
            nets.append(torch.cat(parts))
        w = torch.stack(nets).to(device)
        assert w.shape[1] == L.lib().msched_dqn_param_count(n_in, n_actions)
        self.policy = torch.nn.Parameter(w.clone(), requires_grad=False)
        self.target = w.clone()
        # torch.optim.Adam with its defaults per net (src/Agent.py:313-320), through msched_adam_step
        self.lr, self.opt_step = 1e-3, 0
        self._m, self._v, self._grad = torch.zeros_like(w), torch.zeros_like(w), torch.zeros_like(w)
        self._loss = torch.zeros(units, device=device)
        self.memory = ReplayMemory(memory_size, units, n_in, device)
        self.step_no = 0

    def epsilon(self, round_):
        return self.RUN_END + (self.RUN_START - self.RUN_END) * math.exp(-1.0 * round_ / self.RUN_DECAY)

    def selectAction(self, x, x_stride, env_stride, n_envs, round_, seed, action_rec=None, action_rec_stride=0,
                     u=None, want_q=False, random_policy=False, row_offset=0):
        """DQNEntity.selectAction for every (env, unit).  Returns int32 actions [n_envs, units]."""
        dev = x.device
        M = n_envs * self.units
        action = torch.empty(M, dtype=torch.int32, device=dev)
        q = torch.empty((M, self.A), dtype=torch.float32, device=dev) if want_q else None
        w = self.policy.detach().contiguous()
        desc = L.MschedMlpGroup(self.n_in, self.H, self.A, self.units, 1, 0, w.data_ptr())
        io = L.MschedActorIO()
        io.x, io.x_stride, io.units, io.n_envs = x.data_ptr(), x_stride, self.units, n_envs
        io.env_stride, io.seed, io.step = env_stride, seed, self.step_no
        io.row_offset = row_offset  # global row index of (env 0, unit 0): draws do not depend on the env sharding
        if u is not None:
            u = torch.as_tensor(u, dtype=torch.float32).to(dev).contiguous()
            io.u_override = u.data_ptr()
        io.action = action.data_ptr()
        io.action_rec = None if action_rec is None else action_rec.data_ptr()
        io.action_rec_stride = action_rec_stride
        eps = 1.0 if random_policy else self.epsilon(round_)
        L.check(L.lib().msched_dqn_select(C.byref(desc), C.byref(io), C.c_float(eps),
                                          None if q is None else q.data_ptr(), _stream(dev)))
        self.step_no += 1
        self._keep = (w, u)
        a = action.view(n_envs, self.units)
        return (a, q.view(n_envs, self.units, self.A)) if want_q else a

    def optimize_batch(self, s, a, s2, r):
        """One optimize_model step (src/DQNmodules.py:97-154) for every unit on an explicit batch: s / s2 int16
        [batch, units, n_in], a [batch, units], r [batch, units].  Returns the per-unit loss (device tensor)."""
        from . import policy as P
        batch = s.shape[0]
        s, s2 = s.contiguous(), s2.contiguous()
        a32, r32 = a.to(torch.int32).contiguous(), r.to(torch.float32).contiguous()
        b = L.MschedDqnBatch()
        b.policy, b.target = self.policy.data_ptr(), self.target.data_ptr()
        b.n_in, b.n_hidden, b.n_actions, b.n_nets, b.batch = self.n_in, self.H, self.A, self.units, batch
        b.state, b.next_state, b.action, b.reward = s.data_ptr(), s2.data_ptr(), a32.data_ptr(), r32.data_ptr()
        b.gamma = float(self.gamma)
        b.grad, b.loss = self._grad.data_ptr(), self._loss.data_ptr()
        L.check(L.lib().msched_dqn_grad(C.byref(b), _stream(self.device)))
        self.opt_step += 1
        P.adam_step(self.policy.data.view(-1), self._grad.view(-1), self._m.view(-1), self._v.view(-1), self.lr, self.opt_step)
        self._keep_batch = (s, s2, a32, r32)
        return self._loss

    def optimize_model(self, batch_size, generator=None):
        """optimize_model (src/DQNmodules.py:97-154) for all units at once: sample, gradient kernel, Adam."""
        if len(self.memory) < batch_size:
            return None
        s, a, s2, r = self.memory.sample(batch_size, generator)
        return self.optimize_batch(s, a, s2, r)

    def update_target(self):
        self.target.copy_(self.policy.detach())


class DividedFixPriceDQNAgents:
    """All N agents of src/Agent.py:303-356: N*C acceptor Q-nets and N*L offer Q-nets with their
    target nets, Adam optimizers and replay memories."""

    def __init__(self, world, env):
        self.world, self.env = world, env
        N, C, L_ = world.numberOfAgents, world.numberOfCores, world.collectionLength
        NL = N * L_
        dev = env.core.device
        a = dict(run_start=env.RUN_START, run_end=env.RUN_END, run_decay=env.RUN_DECAY,
                 memory_size=env.REPLAY_MEMORY_SIZE, device=dev)
        self.acceptor = BatchedDQN(3 + 2 * NL, NL + 1, N * C, gamma=env.ACCEPTOR_GAMMA, seed=1, **a)
        self.offer = BatchedDQN(2 * C + 2, C + 1, N * L_, gamma=env.OFFER_GAMMA, seed=2, **a)

    def getActions(self, offerObs, acceptorObs):
        c, lay = self.env.core, self.env.core.layout
        rnd, seed = self.world.round, self.world.seed
        rp = bool(self.world.randomPolicy)
        off = self.world.envOffset
        self.offer.selectAction(offerObs, lay.o_off_row, lay.obs_halfs, c.B, rnd, seed * 2 + 1,
                                action_rec=c.offer_core_actions, action_rec_stride=lay.action_halfs, random_policy=rp,
                                row_offset=off * self.offer.units)
        self.acceptor.selectAction(acceptorObs, lay.o_acc_row, lay.obs_halfs, c.B, rnd, seed * 2,
                                   action_rec=c.acceptor_actions, action_rec_stride=lay.action_halfs, random_policy=rp,
                                   row_offset=off * self.acceptor.units)
        return c.acceptor_actions, c.offer_core_actions

    def updateTargetNets(self):
        self.acceptor.update_target()
        self.offer.update_target()
