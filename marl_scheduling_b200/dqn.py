"""Batched DQN units: the interface of reference src/DQNmodules.py
(`ReplayMemory.push/sample`, `DQNEntity.selectAction`, `optimize_model`) and of
`DividedFixPriceDQNAgent` (src/Agent.py:303-356) with a leading environment dimension.

Rollout: the epsilon-greedy Q-forward of every (environment, unit) runs in the CUDA kernel behind
msched_dqn_select, straight on the observation record, writing the action record.  Learning
(`optimize_model`: SmoothL1Loss against the target net, gradients clamped to [-1, 1], Adam) uses
PyTorch autograd over device-resident replay memories, like the PPO update (N1).
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
        self.policy = torch.nn.Parameter(w.clone())
        self.target = w.clone()
        self.optimizer = torch.optim.Adam([self.policy])
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

    def _forward(self, flat, x):
        """flat [units, pc], x [units, M, in] float -> Q [units, M, A]."""
        n, H, nin, A = flat.shape[0], self.H, self.n_in, self.A
        o = 0
        W1 = flat[:, o:o + H * nin].view(n, H, nin); o += H * nin
        b1 = flat[:, o:o + H]; o += H
        W2 = flat[:, o:o + A * H].view(n, A, H); o += A * H
        b2 = flat[:, o:o + A]
        h = torch.tanh(torch.baddbmm(b1.unsqueeze(1), x, W1.transpose(1, 2)))
        return torch.baddbmm(b2.unsqueeze(1), h, W2.transpose(1, 2))

    def optimize_model(self, batch_size, generator=None):
        """optimize_model (src/DQNmodules.py:97-154) for all units at once."""
        if len(self.memory) < batch_size:
            return None
        s, a, s2, r = self.memory.sample(batch_size, generator)
        x = s.float().permute(1, 0, 2)                      # [units, batch, in]
        x2 = s2.float().permute(1, 0, 2)
        qsa = self._forward(self.policy, x).gather(2, a.t().unsqueeze(-1)).squeeze(-1)
        with torch.no_grad():
            nxt = self._forward(self.target, x2).max(2)[0]
        expected = nxt * self.gamma + r.t()
        loss = torch.nn.functional.smooth_l1_loss(qsa, expected, reduction="none").mean(1).sum()
        self.optimizer.zero_grad()
        loss.backward()
        self.policy.grad.data.clamp_(-1, 1)
        self.optimizer.step()
        return float(loss.detach())

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
