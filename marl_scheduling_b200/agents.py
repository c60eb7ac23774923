"""Batched PPO agents: the reference's per-unit PPO objects (src/PPOmodules.py:75-174, :273-332,
:335-449, :490-597 and the agents' getActions/saveRewards/updateParts in src/Agent.py:495-619,
:669-735) with a leading environment dimension.

Rollout (the hot path): actor forward + categorical sample + log-prob run in the CUDA actor kernel
(msched_actor_forward) straight on the observation record; rewards are routed into device
buffers; returns use msched_returns.  The PPO update itself (clipped surrogate, 0.5*MSE,
-0.01*entropy, two-LR Adam, K epochs, policy_old sync) is SURVEY.md section 8(f) row N1: for the
16-wide nets (divided, locally and globally shared agents) every epoch is ONE forward+backward
kernel (msched_ppo_grad, tensor-core sample reduction) plus msched_adam_step; the aggregated heads
(32/64 hidden neurons, thousands of actions) use PyTorch autograd over the batched buffers.
MSCHED_PPO_UPDATE=autograd forces the autograd path.
"""
from __future__ import annotations

import os
import random

import torch

from . import policy as P
from .distributed import allreduce_gradients, allreduce_mean_

UPDATE_CHUNK = 1 << 20  # samples per autograd chunk (gradient accumulation bounds memory)


class FlatAdam:
    """torch.optim.Adam semantics (src/PPOmodules.py:100-105: one learning rate per head, defaults
    otherwise) over stacked per-net parameter rows through msched_adam_step.  Every net keeps its own
    step count, like the reference's one optimizer per PPO object."""

    def __init__(self, params, lrs):
        self.params, self.lrs = params, lrs
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = [0] * params[0].shape[0]

    def step(self, grads, nets):
        nets = list(nets)
        whole = len(nets) == len(self.t) and len(set(self.t)) == 1
        for n in nets:
            self.t[n] += 1
        for p, g, m, v, lr in zip(self.params, grads, self.m, self.v, self.lrs):
            if whole:
                P.adam_step(p.view(-1), g.view(-1), m.view(-1), v.view(-1), lr, self.t[0])
            else:
                for n in nets:
                    P.adam_step(p[n], g[n], m[n], v[n], lr, self.t[n])

    def state_dict(self):
        return {"kind": "FlatAdam", "t": list(self.t), "m": [x.cpu() for x in self.m], "v": [x.cpu() for x in self.v],
                "lrs": list(self.lrs)}

    def load_state_dict(self, d):
        if d.get("kind") != "FlatAdam":
            raise ValueError("optimizer state is not a FlatAdam state")
        self.t = list(d["t"])
        for dst, src in zip(self.m + self.v, list(d["m"]) + list(d["v"])):
            dst.copy_(src.to(dst.device))


class BatchedPPO:
    """n_nets ActorCritic pairs of one shape (src/PPOmodules.py:25-72), parameters stacked.

    unit u of an environment is served by net (u // unit_div) % n_nets."""

    def __init__(self, n_in, n_actions, n_hidden, n_nets, units, unit_div, lr_actor, lr_critic, gamma,
                 eps_clip, k_epochs, device, seed=0):
        self.n_in, self.A, self.H = n_in, n_actions, n_hidden
        self.n_nets, self.units, self.unit_div = n_nets, units, unit_div
        self.gamma, self.eps_clip, self.K_epochs = gamma, eps_clip, k_epochs
        self.device = device
        init = P.MlpGroup.random(n_in, n_hidden, n_actions, n_nets, device, seed=seed)
        crit = P.MlpGroup.random(n_in, n_hidden, 1, n_nets, device, seed=seed + 7919)
        self.actor = torch.nn.Parameter(init.weights.clone())      # [n_nets, pc_actor]
        self.critic = torch.nn.Parameter(crit.weights.clone())     # [n_nets, pc_critic]
        self.use_kernels = (P.ppo_grad_supported(n_in, n_hidden, n_actions)
                            and os.environ.get("MSCHED_PPO_UPDATE", "kernel") != "autograd")
        if self.use_kernels:
            self.optimizer = FlatAdam([self.actor.data, self.critic.data], [lr_actor, lr_critic])
            # one flat gradient buffer for both heads: a single all-reduce per epoch under data parallelism
            self._grad = torch.zeros(self.actor.numel() + self.critic.numel(), device=device)
            self._ga = self._grad[: self.actor.numel()].view_as(self.actor)
            self._gc = self._grad[self.actor.numel():].view_as(self.critic)
            self._ws, self._ids = None, {}
        else:
            self.optimizer = torch.optim.Adam([{"params": [self.actor], "lr": lr_actor},
                                               {"params": [self.critic], "lr": lr_critic}])
        self.policy_old = P.MlpGroup(n_in, n_hidden, n_actions, self.actor.detach().clone(), device,
                                     unit_div=unit_div)
        self.buf_x, self.buf_a, self.buf_lp, self.buf_r = [], [], [], []
        self.step_no = 0
        self.profile = None  # a list: every kernel-path epoch appends its (start, grad, all-reduce, Adam) CUDA events

    # -- experience buffer slots (the one-launch policy step writes state / action / log-prob in place) --------
    def slot(self, n_envs, x_stride):
        """Views of the next experience-buffer slot for msched_policy_step: (x int16 [B,U,x_stride], action int32
        [B,U], logprob float32 [B,U]).  The buffer is a ring of whole steps that doubles when it is full."""
        t = len(self.buf_a)
        ring = getattr(self, "_ring", None)
        if ring is None or ring["x"].shape[1] != n_envs or ring["x"].shape[3] != x_stride or t >= ring["x"].shape[0]:
            cap = 32 if ring is None else 2 * ring["x"].shape[0]
            dev = self.device
            new = dict(x=torch.zeros((cap, n_envs, self.units, x_stride), dtype=torch.int16, device=dev),
                       a=torch.zeros((cap, n_envs, self.units), dtype=torch.int32, device=dev),
                       lp=torch.zeros((cap, n_envs, self.units), dtype=torch.float32, device=dev))
            self._ring = ring = new  # earlier steps keep referring to the old ring's tensors
        return ring["x"][t], ring["a"][t], ring["lp"][t]

    def commit_slot(self, x_slot, a_slot, lp_slot, lead):
        """Record a slot the kernel has filled: buffer.states / actions / logprobs of src/PPOmodules.py:114-125."""
        self.buf_x.append(x_slot[:, :, lead: lead + self.n_in])
        self.buf_a.append(a_slot)
        self.buf_lp.append(lp_slot)
        self.step_no += 1

    # -- rollout ---------------------------------------------------------------------------------
    def selectAction(self, x, x_stride, env_stride, n_envs, seed, action_rec=None, action_rec_stride=0,
                     gather_core=None, n_cores=0, env_offset=0):
        """PPO.selectAction for every (env, unit): x is an int16 view whose rows are the units'
        observations.  Returns actions int32 [n_envs, units]; stores state/action/logprob.
        action_rec: int16 view into the environment's action record that the kernel fills
        directly; gather_core: core chooser actions (price chooser launch, see policy.actor_forward);
        env_offset: global index of this shard's env 0 (the sampling draws are keyed on the GLOBAL row, so a
        sharded rollout draws what the unsharded one would)."""
        x_used = None
        if gather_core is not None:
            x_used = torch.empty((n_envs, self.units, self.n_in), dtype=torch.int16, device=x.device)
        act, lp, _ = P.actor_forward(self.policy_old, x, x_stride, self.units, n_envs,
                                     env_stride=env_stride, seed=seed, step=self.step_no,
                                     row_offset=env_offset * self.units,
                                     action_rec=action_rec, action_rec_stride=action_rec_stride,
                                     gather_core=gather_core, n_cores=n_cores, x_used=x_used)
        self.step_no += 1
        # the PPO buffer keeps the observation like buffer.states
        self.buf_x.append(x_used if x_used is not None else x.reshape(n_envs, self.units, self.n_in).clone())
        self.buf_a.append(act.view(n_envs, self.units))
        self.buf_lp.append(lp.view(n_envs, self.units))
        return act.view(n_envs, self.units)

    def saveReward(self, r):
        # a COPY: the reward tensors env.step returns are views into the result record, which a later step
        # overwrites (float32 planes would otherwise be stored by reference)
        self.buf_r.append(r.reshape(r.shape[0], self.units).to(torch.float32, copy=True))

    # -- update (N1) -------------------------------------------------------------------------------
    def _forward(self, flat, x, A):
        """flat [n, pc], x [n, M, in] -> [n, M, A] (Linear-Tanh-Linear-Tanh-Linear)."""
        n, H, nin = flat.shape[0], self.H, self.n_in
        o = 0
        W1 = flat[:, o:o + H * nin].view(n, H, nin); o += H * nin
        b1 = flat[:, o:o + H]; o += H
        W2 = flat[:, o:o + H * H].view(n, H, H); o += H * H
        b2 = flat[:, o:o + H]; o += H
        W3 = flat[:, o:o + A * H].view(n, A, H); o += A * H
        b3 = flat[:, o:o + A]
        h = torch.tanh(torch.baddbmm(b1.unsqueeze(1), x, W1.transpose(1, 2)))
        h = torch.tanh(torch.baddbmm(b2.unsqueeze(1), h, W2.transpose(1, 2)))
        return torch.baddbmm(b3.unsqueeze(1), h, W3.transpose(1, 2))

    def update(self, unit_subset=None):
        """PPO.update (src/PPOmodules.py:127-174) for all nets at once.  Returns are normalised per
        (env, unit) over the buffer like one reference world would; every net then takes K full-
        batch epochs over the samples of the units it serves (unit_subset restricts them, for the
        shared-parameter sampling rule)."""
        T = len(self.buf_r)
        if T < 2:
            return
        B = self.buf_r[0].shape[0]
        U = self.units
        r = torch.stack(self.buf_r).reshape(T, B * U)
        G = P.returns(r, self.gamma, normalise=True).view(T, B, U)
        units = list(range(U)) if unit_subset is None else list(unit_subset)
        nets = sorted({(u // self.unit_div) % self.n_nets for u in units})
        per_net = {n: [u for u in units if (u // self.unit_div) % self.n_nets == n] for n in nets}
        m = max(len(v) for v in per_net.values())
        if any(len(v) != m for v in per_net.values()):
            raise ValueError("unit subset must give every net the same number of units")
        if self.use_kernels:
            return self._update_kernels(T, B, U, G, nets, per_net)
        X = torch.stack(self.buf_x).float()                       # [T,B,U,in]
        Aold = torch.stack(self.buf_a).long()
        LPold = torch.stack(self.buf_lp)
        idx = torch.tensor([per_net[n] for n in nets], device=self.device)      # [n, m]
        def gather(t):  # [T,B,U,...] -> [n, T*B*m, ...]
            g = t[:, :, idx]                                          # [T,B,n,m,...]
            g = g.permute(2, 0, 1, 3, *range(4, g.dim())).contiguous()
            return g.view(len(nets), T * B * m, *g.shape[4:])
        Xn, An, LPn, Gn = gather(X), gather(Aold), gather(LPold), gather(G)
        M = Xn.shape[1]
        sel = torch.tensor(nets, device=self.device)
        for _ in range(self.K_epochs):
            self.optimizer.zero_grad()
            # MSELoss is a mean over the whole batch of a net: first pass for the values
            with torch.no_grad():
                mse = torch.zeros(len(nets), device=self.device)
                for c0 in range(0, M, UPDATE_CHUNK):
                    v = self._forward(self.critic[sel], Xn[:, c0:c0 + UPDATE_CHUNK], 1).squeeze(-1)
                    mse += ((v - Gn[:, c0:c0 + UPDATE_CHUNK]) ** 2).sum(1)
                mse /= M
            for c0 in range(0, M, UPDATE_CHUNK):
                xs = Xn[:, c0:c0 + UPDATE_CHUNK]
                logits = self._forward(self.actor[sel], xs, self.A)
                logp_all = torch.log_softmax(logits, -1)
                logp = logp_all.gather(-1, An[:, c0:c0 + UPDATE_CHUNK].unsqueeze(-1)).squeeze(-1)
                ent = -(logp_all.exp() * logp_all).sum(-1)
                v = self._forward(self.critic[sel], xs, 1).squeeze(-1)
                g = Gn[:, c0:c0 + UPDATE_CHUNK]
                ratios = torch.exp(logp - LPn[:, c0:c0 + UPDATE_CHUNK])
                adv = g - v.detach()
                surr = torch.min(ratios * adv, torch.clamp(ratios, 1 - self.eps_clip, 1 + self.eps_clip) * adv)
                # d/dtheta of mean(-surr + 0.5*MSE - 0.01*H): the MSE term is a batch mean added
                # to every element, so its gradient is that of 0.5*mean((v-g)^2)
                loss = (-surr - 0.01 * ent).sum(1) / M + 0.5 * ((v - g) ** 2).sum(1) / M
                loss.sum().backward()
            # data-parallel env shards: one flat-bucket all-reduce per epoch (NCCL over NVLink)
            allreduce_gradients([self.actor, self.critic])
            self.optimizer.step()
        return float(mse.mean())

    def _update_kernels(self, T, B, U, G, nets, per_net):
        """K epochs of msched_ppo_grad (forward + backward of every selected net in one launch) and
        msched_adam_step on the stacked buffers; the observations stay int16."""
        dev = self.device
        X = torch.stack(self.buf_x).view(T * B, U, self.n_in)       # int16 [TB,U,in]
        Aold = torch.stack(self.buf_a).view(T * B, U)                 # int32
        LPold = torch.stack(self.buf_lp).view(T * B, U)
        key = tuple((n, tuple(per_net[n])) for n in nets)
        if key not in self._ids:
            self._ids[key] = (torch.tensor(nets, dtype=torch.int32, device=dev),
                              torch.tensor([per_net[n] for n in nets], dtype=torch.int32, device=dev))
        net_ids, unit_ids = self._ids[key]
        stats = None
        ev = None
        for _ in range(self.K_epochs):
            if self.profile is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
            if len(nets) != self.n_nets:
                self._grad.zero_()
            stats, self._ws = P.ppo_grad(self.actor.data, self.critic.data, self.n_in, self.A, X, Aold, LPold,
                                         G.view(T * B, U), net_ids, unit_ids, self._ga, self._gc,
                                         eps_clip=self.eps_clip, stats=stats, workspace=self._ws)
            if ev:
                ev[1].record()
            # data-parallel env shards: one flat-bucket all-reduce per epoch (NCCL over NVLink)
            allreduce_mean_(self._grad)
            if ev:
                ev[2].record()
            self.optimizer.step([self._ga, self._gc], nets)
            if ev:
                ev[3].record()
                self.profile.append(ev)
        return stats[:, 1].mean()

    def sync_old_and_clear(self):
        self.policy_old.weights.copy_(self.actor.detach())
        self.buf_x, self.buf_a, self.buf_lp, self.buf_r = [], [], [], []


def _input_bound(world):
    """Largest |value| an observation can hold (priorities, remaining lengths, prices, the -1/-2/-5 markers): lets
    msched_policy_step take the inputs as exact fp16 operands (tensor-core kernel) when it is at most 511."""
    vals = [5] + [abs(int(v)) for v in world.possibleJobPriorities] + [abs(int(v)) for v in world.possibleJobLengths]
    if not world.freePrices:
        vals += [abs(int(v)) for v in world.listOfFixPrices]
    return max(vals)


def _one_launch_ok(core, acceptorObs, offerObs, env_offset, acceptor, offer, price=None):
    """msched_policy_step applies when the nets have one of its shapes, the shard starts on an even global env
    and the observations handed in are the views of the env's current observation record (not copies)."""
    if os.environ.get("MSCHED_POLICY_STEP", "1") == "0" or (env_offset & 1):
        return False
    if not P.policy_step_supported(acceptor.policy_old, offer.policy_old, None if price is None else price.policy_old):
        return False
    base, lay = core._obs_buffer().data_ptr(), core.layout
    return (acceptorObs.data_ptr() == base + 2 * lay.o_acceptor and offerObs.data_ptr() == base + 2 * lay.o_offer)


class DividedFixedPricePPOAgents:
    """All N agents of src/Agent.py:495-536 (divided), :539-576 (globally shared) or :669-735
    (locally shared) at once: N*C acceptor units and N*L offer units."""

    def __init__(self, world, env, sharing=None):
        self.world, self.env, self.sharing = world, env, sharing
        N, C, L = world.numberOfAgents, world.numberOfCores, world.collectionLength
        NL = N * L
        dev = env.core.device
        if sharing == "global":
            na, no, da, do = 1, 1, 1, 1
        elif sharing == "local":
            na, no, da, do = N, N, C, L
        else:
            na, no, da, do = N * C, N * L, 1, 1
        self.acceptor = BatchedPPO(3 + 2 * NL, NL + 1, 16, na, N * C, da, env.LR_ACTOR, env.LR_CRITIC,
                                   env.ACCEPTOR_GAMMA, env.EPS_CLIP, env.ACCEPTOR_K_EPOCHS, dev, seed=1)
        self.offer = BatchedPPO(2 * C + 2, C + 1, 16, no, N * L, do, env.LR_ACTOR, env.LR_CRITIC,
                                env.OFFER_GAMMA, env.EPS_CLIP, env.OFFER_K_EPOCHS, dev, seed=2)
        self.CENTRALISATION_SAMPLE = env.CENTRALISATION_SAMPLE
        # which units a shared net learns from (src/SchedulingEnvironment.py:314-329, src/Agent.py:708-728): a
        # private generator seeded like the world, so every data-parallel rank samples the same units
        self._rng = random.Random(world.seed)

    def getActions(self, offerObs, acceptorObs):
        c, lay = self.env.core, self.env.core.layout
        B, N, C, L = c.B, c.N, c.C, c.Lc
        seed = self.world.seed
        # the kernels write the chosen actions straight into the env's action record
        eo = self.world.envOffset
        if _one_launch_ok(c, acceptorObs, offerObs, eo, self.acceptor, self.offer):
            # every unit of the step in ONE launch (msched_policy_step); state / action / log-prob written in place
            sa, so = self.acceptor.slot(B, lay.o_acc_row), self.offer.slot(B, lay.o_off_row)
            ga = P.policy_step_group(self.acceptor.policy_old, N * C, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor,
                                     seed * 2, sa[1], sa[2], x_used=sa[0])
            go = P.policy_step_group(self.offer.policy_old, N * L, lay.o_offer, lay.o_off_row, lay.a_offer_core,
                                     seed * 2 + 1, so[1], so[2], x_used=so[0])
            P.policy_step(c._obs_buffer(), lay.obs_halfs, B, C, ga, go, None, action_rec=c.action,
                          action_rec_stride=lay.action_halfs, env_offset=eo, step=self.acceptor.step_no,
                          input_bound=_input_bound(self.world))
            self.acceptor.commit_slot(*sa, lay.o_acceptor & 1)
            self.offer.commit_slot(*so, 0)
            return c.acceptor_actions, c.offer_core_actions
        self.offer.selectAction(offerObs, lay.o_off_row, lay.obs_halfs, B, seed * 2 + 1,
                                action_rec=c.offer_core_actions, action_rec_stride=lay.action_halfs, env_offset=eo)
        self.acceptor.selectAction(acceptorObs, lay.o_acc_row, lay.obs_halfs, B, seed * 2,
                                   action_rec=c.acceptor_actions, action_rec_stride=lay.action_halfs, env_offset=eo)
        # SchedulingEnv.getActionForAllAgents returns (acceptorActions, offerActions)
        return c.acceptor_actions, c.offer_core_actions

    def saveRewards(self, offerRewards, acceptorRewards, agentReward):
        self.offer.saveReward(offerRewards)
        self.acceptor.saveReward(acceptorRewards)

    def updateParts(self):
        for ppo, sub in ((self.acceptor, self.world.numberOfCores), (self.offer, self.world.collectionLength)):
            if self.sharing is None:
                ppo.update()
            elif self.sharing == "global":   # src/SchedulingEnvironment.py:314-329
                for _ in range(self.CENTRALISATION_SAMPLE):
                    ppo.update([self._rng.randint(0, ppo.units - 1)])
            else:                            # src/Agent.py:708-728: every agent draws its own sub-unit
                N = self.world.numberOfAgents
                for _ in range(self.CENTRALISATION_SAMPLE):
                    ppo.update([a * sub + self._rng.randint(0, sub - 1) for a in range(N)])
            ppo.sync_old_and_clear()


class DividedFreePricePPOAgents:
    """src/Agent.py:579-619 + FreePriceOfferPPO (src/PPOmodules.py:273-332): acceptor units as
    above; every offer unit is a core chooser followed by a price chooser whose input is
    [core prio, core rem, slot prio, slot rem] of the chosen core (quirk Q1: action 0 means "no
    offer" to the chooser, a dummy [-5]*4 input is pushed through the price net and price -5 is
    reported, while the world still maps action 0 to core 1)."""

    def __init__(self, world, env):
        self.world, self.env = world, env
        N, C, L = world.numberOfAgents, world.numberOfCores, world.collectionLength
        NL = N * L
        dev = env.core.device
        a = dict(lr_actor=env.LR_ACTOR, lr_critic=env.LR_CRITIC, eps_clip=env.EPS_CLIP, device=dev)
        self.acceptor = BatchedPPO(3 + 2 * NL, NL + 1, 16, N * C, N * C, 1, gamma=env.ACCEPTOR_GAMMA,
                                   k_epochs=env.ACCEPTOR_K_EPOCHS, seed=1, **a)
        self.core = BatchedPPO(2 * C + 2, C + 1, 16, NL, NL, 1, gamma=env.OFFER_GAMMA,
                               k_epochs=env.RAW_K_EPOCHS, seed=2, **a)
        self.price = BatchedPPO(4, world.maxSumToOffer + 1, 16, NL, NL, 1, gamma=env.OFFER_GAMMA,
                                k_epochs=env.RAW_K_EPOCHS, seed=3, **a)

    def getActions(self, offerObs, acceptorObs):
        c, lay = self.env.core, self.env.core.layout
        B, N, C, L = c.B, c.N, c.C, c.Lc
        seed = self.world.seed
        if (_one_launch_ok(c, acceptorObs, offerObs, self.world.envOffset, self.acceptor, self.core, self.price)
                and self.core.step_no == self.price.step_no == self.acceptor.step_no):
            # acceptor units + core chooser + price chooser of every offer unit in ONE launch (msched_policy_step)
            NL = N * L
            sa, sc, sp = self.acceptor.slot(B, lay.o_acc_row), self.core.slot(B, lay.o_off_row), self.price.slot(B, 4)
            ga = P.policy_step_group(self.acceptor.policy_old, N * C, lay.o_acceptor, lay.o_acc_row, lay.a_acceptor,
                                     seed * 3, sa[1], sa[2], x_used=sa[0])
            gc = P.policy_step_group(self.core.policy_old, NL, lay.o_offer, lay.o_off_row, lay.a_offer_core,
                                     seed * 3 + 1, sc[1], sc[2], x_used=sc[0])
            gp = P.policy_step_group(self.price.policy_old, NL, lay.o_offer, lay.o_off_row, lay.a_offer_price,
                                     seed * 3 + 2, sp[1], sp[2], x_used=sp[0])
            P.policy_step(c._obs_buffer(), lay.obs_halfs, B, C, ga, gc, gp, action_rec=c.action,
                          action_rec_stride=lay.action_halfs, env_offset=self.world.envOffset, step=self.acceptor.step_no,
                          input_bound=_input_bound(self.world))
            self.acceptor.commit_slot(*sa, lay.o_acceptor & 1)
            self.core.commit_slot(*sc, 0)
            self.price.commit_slot(*sp, 0)
            return c.acceptor_actions, (c.offer_core_actions, c.offer_price_actions)
        eo = self.world.envOffset
        core = self.core.selectAction(offerObs, lay.o_off_row, lay.obs_halfs, B, seed * 3 + 1,
                                      action_rec=c.offer_core_actions, action_rec_stride=lay.action_halfs, env_offset=eo)
        # price chooser: the kernel slices [core prio, core rem, slot prio, slot rem] of the chosen
        # core out of the offer observation row ([-5]*4 and reported price -5 for core action 0)
        self.price.selectAction(offerObs, lay.o_off_row, lay.obs_halfs, B, seed * 3 + 2,
                                action_rec=c.offer_price_actions, action_rec_stride=lay.action_halfs,
                                gather_core=core.reshape(-1), n_cores=C, env_offset=eo)
        self.acceptor.selectAction(acceptorObs, lay.o_acc_row, lay.obs_halfs, B, seed * 3,
                                   action_rec=c.acceptor_actions, action_rec_stride=lay.action_halfs, env_offset=eo)
        return c.acceptor_actions, (c.offer_core_actions, c.offer_price_actions)

    def saveRewards(self, offerRewards, acceptorRewards, agentReward):
        coreChooser, priceChooser = offerRewards
        self.core.saveReward(coreChooser)
        self.price.saveReward(priceChooser)
        self.acceptor.saveReward(acceptorRewards)

    def updateParts(self):
        for ppo in (self.acceptor, self.core, self.price):
            ppo.update()
            ppo.sync_old_and_clear()


def _digits(number, base, dimensionality):
    """numberToNDimensionalAction (src/Agent.py:644-666) for tensors: digit k = (n // base**k) % base,
    index 0 = least significant."""
    return torch.stack([(number // (base ** k)) % base for k in range(dimensionality)], dim=-1)


class AggregatedFixPricePPOAgents:
    """All N agents of src/Agent.py:359-390 (semi-aggregated): per agent one AggregatedAcceptorPPO
    (in C(3+2NL), (NL+1)^C actions, 32 neurons) and one AggregatedOfferPPO (in 2C+2L, (C+1)^L actions)
    (src/PPOmodules.py:177-211).  The heads run on the tensor cores (msched_actor_forward tiles the last
    layer over the actions); the aggregated action number is decoded into the per-core / per-slot
    actions the world consumes."""

    def __init__(self, world, env):
        self.world, self.env = world, env
        N, C, L = world.numberOfAgents, world.numberOfCores, world.collectionLength
        NL = N * L
        dev = env.core.device
        if (NL + 1) ** C > 32767 or (C + 1) ** L > 32767:
            raise ValueError("aggregated action space too large (more than 32,767 actions per unit)")
        a = dict(lr_actor=env.LR_ACTOR, lr_critic=env.LR_CRITIC, eps_clip=env.EPS_CLIP, device=dev)
        self.acceptor = BatchedPPO(C * (3 + 2 * NL), (NL + 1) ** C, 32, N, N, 1, gamma=env.ACCEPTOR_GAMMA,
                                   k_epochs=env.ACCEPTOR_K_EPOCHS, seed=1, **a)
        self.offer = BatchedPPO(2 * C + 2 * L, (C + 1) ** L, 32, N, N, 1, gamma=env.OFFER_GAMMA,
                                k_epochs=env.OFFER_K_EPOCHS, seed=2, **a)

    def getActions(self, offerObs, acceptorObs):
        """offerObs int16 [B,N,2C+2L], acceptorObs int16 [B,N,C(3+2NL)] (env.aggregatedObservations)."""
        c = self.env.core
        B, N, C, L = c.B, c.N, c.C, c.Lc
        seed = self.world.seed
        eo = self.world.envOffset
        acc = self.acceptor.selectAction(acceptorObs.contiguous(), acceptorObs.shape[-1], 0, B, seed * 2, env_offset=eo)
        off = self.offer.selectAction(offerObs.contiguous(), offerObs.shape[-1], 0, B, seed * 2 + 1, env_offset=eo)
        nd_acc = _digits(acc.long(), c.NL + 1, C)      # [B,N,C]
        nd_off = _digits(off.long(), C + 1, L)         # [B,N,L]
        return nd_acc, nd_off

    def saveRewards(self, offerUnitRewards, acceptorUnitRewards, agentReward):
        # src/SchedulingEnvironment.py:223-225, src/Agent.py:388-390
        self.acceptor.saveReward(agentReward)
        self.offer.saveReward(offerUnitRewards[..., 0])

    def updateParts(self):
        for ppo in (self.acceptor, self.offer):
            ppo.update()
            ppo.sync_old_and_clear()


class FullyAggregatedFixPricePPOAgents:
    """src/Agent.py:393-492: one FullyAggregatedPPO per agent (64 neurons) on
    cat(offerObs, acceptorObs); action = acceptor number * (C+1)^L + offer number."""

    def __init__(self, world, env):
        self.world, self.env = world, env
        N, C, L = world.numberOfAgents, world.numberOfCores, world.collectionLength
        NL = N * L
        self.divisor = (C + 1) ** L
        n_actions = ((NL + 1) ** C) * self.divisor
        if n_actions > 32767:
            raise ValueError("fully aggregated action space too large (more than 32,767 actions)")
        self.unit = BatchedPPO(C * (3 + 2 * NL) + 2 * C + 2 * L, n_actions, 64, N, N, 1, env.LR_ACTOR,
                               env.LR_CRITIC, env.ACCEPTOR_GAMMA, env.EPS_CLIP, env.ACCEPTOR_K_EPOCHS,
                               env.core.device, seed=1)

    def getActions(self, offerObs, acceptorObs):
        c = self.env.core
        x = torch.cat([offerObs, acceptorObs], dim=-1).contiguous()   # src/Agent.py:463-467
        n = self.unit.selectAction(x, x.shape[-1], 0, c.B, self.world.seed, env_offset=self.world.envOffset).long()
        nd_acc = _digits(n // self.divisor, c.NL + 1, c.C)
        nd_off = _digits(n % self.divisor, c.C + 1, c.Lc)
        return nd_acc, nd_off

    def saveRewards(self, offerUnitRewards, acceptorUnitRewards, agentReward):
        # src/SchedulingEnvironment.py:243-247: agentReward[i] + offerUnitRewards[i][0]
        self.unit.saveReward(agentReward.float() + offerUnitRewards[..., 0].float())

    def updateParts(self):
        self.unit.update()
        self.unit.sync_old_and_clear()
