"""BatchedSchedulingEnv: B independent reference worlds advanced by one CUDA launch per step.

This is the device-resident core that the drop-in classes in SchedulingEnvironment.py wrap.
torch is used for device memory and streams only; all compute goes through the C-ABI
(include/msched.h) into the sm_100a kernels.  Argument order and meaning of `step` mirror
SchedulingEnv.step(offerActions, acceptorActions, auctioneer_action)
(reference src/SchedulingEnvironment.py:32-83) with a leading env dimension.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def world_params_from_dom(dom, free):
    """Small helper for tests/bench: compact domain dict -> reference World(params) keys."""
    return dict(
        freePrices=bool(free), fixPricesList=list(dom.get("fix", [])),
        numberOfAgents=dom["N"], numberOfCores=dom["C"], collectionLength=dom["L"],
        possibleJobPriorities=list(dom["prios"]), possibleJobLengths=list(dom["lens"]),
        probabilities=list(dom["probs"]), newJobsPerRoundPerAgent=dom.get("newJobs", 1),
        rewardMultiplier=dom.get("mult", 1), episodeLength=dom.get("episodeLength", 100),
        maxVisibleOffers=4)


class BatchedSchedulingEnv:
    def __init__(self, B, world_params, reward="fix", auction="external", spawn="philox",
                 chain_capacity=64, seed=0, env_offset=0, net_zero_offer_reward=0.5, device=0):
        if not torch.cuda.is_available():
            raise L.MschedError("no CUDA device: marl_scheduling_b200 has no CPU fallback")
        self.lib = L.lib()
        self.cfg = L.make_config(B, world_params, reward, auction, spawn, chain_capacity, seed,
                                 env_offset, net_zero_offer_reward)
        self.layout = lay = L.get_layout(self.cfg)
        self.B, self.N, self.C, self.Lc = B, self.cfg.N, self.cfg.C, self.cfg.L
        self.NL = self.N * self.Lc
        self.newJobs = self.cfg.newJobsPerRound
        self.free = bool(self.cfg.freePrices)
        self.agg = self.cfg.rewardVariant == 3
        self.device = torch.device("cuda", device)
        self.handle = C.c_void_p()
        L.check(self.lib.msched_create(C.byref(self.cfg), device, C.byref(self.handle)))
        Bp = lay.padded_envs
        dev = self.device
        self.state = torch.zeros((Bp, lay.state_words), dtype=torch.int32, device=dev)
        self.chain = torch.zeros((Bp, lay.chain_words), dtype=torch.int32, device=dev)
        self.action = torch.zeros((Bp, lay.action_halfs), dtype=torch.int16, device=dev)
        self.result = torch.zeros((Bp, lay.result_words), dtype=torch.int32, device=dev)
        # step() alternates between two result records and two observation records: what one step returned
        # (views into them) stays valid until the SECOND-next step, so the reference's `old = new` idiom
        # around env.step (src/trainDQN.py:171-186) sees two different tensors
        self._result_ring = [self.result, None]
        self._obs_ring = [None, None]
        self._flip = 0
        self._obs = None
        self._ids = None
        L.check(self.lib.msched_bind_state(self.handle, self.state.data_ptr(), self.chain.data_ptr()))
        self.reset()

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.msched_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def info(self):
        """Which kernels the handle launches: dict(step_impl, fuses_observations, ...)."""
        i = L.MschedInfo()
        L.check(self.lib.msched_get_info(self.handle, C.byref(i)))
        return dict(step_impl=("lane", "coop", "fused", "warp")[i.step_impl],
                    fuses_observations=bool(i.fuses_observations), envs_per_cta=i.envs_per_cta,
                    threads_per_cta=i.threads_per_cta, smem_bytes_per_cta=i.smem_bytes_per_cta)

    @property
    def round(self):
        r = C.c_int64()
        L.check(self.lib.msched_get_round(self.handle, C.byref(r)))
        return int(r.value)

    def set_device_round(self, enable=True):
        """world.round in a device counter (advanced by a one-thread kernel after every step), so
        that a captured step can be replayed from a CUDA graph."""
        L.check(self.lib.msched_set_round_mode(self.handle, int(bool(enable)), self._stream()))

    def reset(self):
        """Fresh worlds (World.__init__, reference src/world.py:210-254)."""
        L.check(self.lib.msched_reset(self.handle, self._stream()))

    # ------------------------------------------------------------------ action record views
    def _aview(self, off, n, shape):
        return self.action[: self.B, off: off + n].view(self.B, *shape) if off >= 0 else None

    @property
    def acceptor_actions(self):  # [B,N,C] int16 view into the action record
        return self._aview(self.layout.a_acceptor, self.N * self.C, (self.N, self.C))

    @property
    def offer_core_actions(self):  # [B,N,L]
        return self._aview(self.layout.a_offer_core, self.NL, (self.N, self.Lc))

    @property
    def offer_price_actions(self):  # [B,N,L] (free prices) or None
        return self._aview(self.layout.a_offer_price, self.NL, (self.N, self.Lc))

    @property
    def auctioneer_actions(self):  # [B,C] (external auctioneer) or None
        return self._aview(self.layout.a_auctioneer, self.C, (self.C,))

    @property
    def spawn_kinds(self):  # [B,N,newJobs] (spawn='kinds') or None
        return self._aview(self.layout.a_spawn_kind, self.N * self.newJobs, (self.N, self.newJobs))

    def set_actions(self, offer_core, acceptor, auctioneer=None, offer_price=None, spawn_kind=None):
        """Copy per-field action tensors (any integer dtype, device or host) into the record."""
        def put(view, src, name):
            if view is None:
                if src is not None:
                    raise ValueError(f"{name} not part of this configuration's action record")
                return
            if src is None:
                raise ValueError(f"{name} required")
            src = torch.as_tensor(src)
            if (src.dtype == view.dtype and src.device == view.device and src.shape == view.shape
                    and src.data_ptr() == view.data_ptr() and src.stride() == view.stride()):
                return  # the policy kernels already wrote this field of the record
            view.copy_(src.to(self.device, non_blocking=True).reshape(view.shape))
        put(self.offer_core_actions, offer_core, "offer_core")
        put(self.acceptor_actions, acceptor, "acceptor")
        put(self.auctioneer_actions, auctioneer, "auctioneer")
        put(self.offer_price_actions, offer_price, "offer_price")
        put(self.spawn_kinds, spawn_kind, "spawn_kind")

    # ------------------------------------------------------------------ the hot path
    def step_records(self, action=None, result=None, spawn_u=None):
        """One step on caller-provided (or the env's own) device records.  Asynchronous."""
        action = self.action if action is None else action
        result = self.result if result is None else result
        su = None
        if spawn_u is not None:
            su = torch.as_tensor(spawn_u, dtype=torch.float64).to(self.device).contiguous()
            self._keep_su = su
        L.check(self.lib.msched_step(self.handle, action.data_ptr(),
                                     None if su is None else su.data_ptr(), result.data_ptr(),
                                     self._stream()))
        return result

    def step_observe_records(self, action=None, result=None, spawn_u=None, obs=None):
        """step_records + the dense observations of the new state in one C-ABI call
        (msched_step_observe: ONE fused launch for the compile-time domains).  Asynchronous."""
        action = self.action if action is None else action
        result = self.result if result is None else result
        obs = self._obs_buffer() if obs is None else obs
        su = None
        if spawn_u is not None:
            su = torch.as_tensor(spawn_u, dtype=torch.float64).to(self.device).contiguous()
            self._keep_su = su
        L.check(self.lib.msched_step_observe(self.handle, action.data_ptr(),
                                             None if su is None else su.data_ptr(), result.data_ptr(),
                                             obs.data_ptr(), self._stream()))
        return result, obs

    def step_multi_records(self, actions, results, obs=None, obs_every=False):
        """msched_step_multi: T = actions.shape[0] consecutive steps in ONE launch (fused-kernel domains).  actions int16
        [T, padded_envs, action_halfs], results int32 [T, padded_envs, result_words]; obs: the env's observation
        record (last step) by default, or int16 [T, padded_envs, obs_halfs] with obs_every.  Asynchronous."""
        T = int(actions.shape[0])
        assert actions.is_contiguous() and results.is_contiguous() and results.shape[0] == T
        if obs is None and not obs_every:
            obs = self._obs_buffer()
        L.check(self.lib.msched_step_multi(self.handle, actions.data_ptr(), T, results.data_ptr(),
                                           None if obs is None else obs.data_ptr(), 1 if obs_every else 0, self._stream()))
        return results, obs

    def step_compact_records(self, action=None, result=None, spawn_u=None, cobs=None):
        """step_records + the COMPACT observations of the new state (msched_step_compact: one launch on the
        warp-per-environment kernel of the large domains).  Asynchronous."""
        action = self.action if action is None else action
        result = self.result if result is None else result
        cobs = self._cobs_buffer() if cobs is None else cobs
        su = None
        if spawn_u is not None:
            su = torch.as_tensor(spawn_u, dtype=torch.float64).to(self.device).contiguous()
            self._keep_su = su
        L.check(self.lib.msched_step_compact(self.handle, action.data_ptr(),
                                             None if su is None else su.data_ptr(), result.data_ptr(),
                                             cobs.data_ptr(), self._stream()))
        return result, cobs

    def step(self, offer_core, acceptor, auctioneer=None, offer_price=None, spawn_kind=None,
             spawn_u=None, observe=False):
        """SchedulingEnv.step(offerActions, acceptorActions, auctioneer_action) over B envs.
        observe=True also refreshes the observation record (see obs_views)."""
        self.set_actions(offer_core, acceptor, auctioneer, offer_price, spawn_kind)
        self._flip ^= 1
        k = self._flip
        if self._result_ring[k] is None:
            self._result_ring[k] = torch.zeros_like(self.result)
        self.result = self._result_ring[k]
        if observe:
            if self._obs_ring[k] is None:
                self._obs_ring[k] = torch.zeros((self.layout.padded_envs, self.layout.obs_halfs), dtype=torch.int16,
                                                device=self.device)
            self._obs = self._obs_ring[k]
            self.step_observe_records(spawn_u=spawn_u)
        else:
            self.step_records(spawn_u=spawn_u)
        return self.rewards()

    def step_host(self, action_host, result_host, observe=False):
        """The C-ABI call with HOST buffers (pinned int16 / int32 tensors): H2D, step, D2H, sync.
        observe=True also refreshes the device-resident observation record (see obs_views)."""
        L.check(self.lib.msched_step_host(self.handle, action_host.data_ptr(), result_host.data_ptr(),
                                          self._obs_buffer().data_ptr() if observe else None, self._stream()))
        return result_host

    def compact_result_layout(self):
        """MschedCompactResultLayout of this configuration (raises MschedError if its rewards are not half-exact)."""
        cl = L.MschedCompactResultLayout()
        L.check(self.lib.msched_get_compact_result_layout(C.byref(self.cfg), C.byref(cl)))
        return cl

    def step_host_compact(self, action_host, cresult_host, observe=False):
        """msched_step_host_compact: pinned int16 action records in, COMPACT result records out (pinned int32
        [B, words]); decode with compact_rewards()."""
        L.check(self.lib.msched_step_host_compact(self.handle, action_host.data_ptr(), cresult_host.data_ptr(),
                                                  self._obs_buffer().data_ptr() if observe else None, self._stream()))
        return cresult_host

    def compact_rewards(self, cresult):
        """Decode compact result records (any device) into the dict rewards() returns (float32 / int16 planes)."""
        cl, B, N, lay = self.compact_result_layout(), self.B, self.N, self.layout
        r = cresult[:B]
        h = r.view(torch.int16)
        out = {}
        f = lambda off, n: h[:, off: off + n].contiguous().view(torch.float16).float()
        out["offer"] = f(cl.c_offer, N * lay.RL).view(B, N, lay.RL)
        out["price"] = f(cl.c_price, N * lay.RL).view(B, N, lay.RL) if cl.c_price >= 0 else None
        out["acceptor"] = h[:, cl.c_acceptor: cl.c_acceptor + N * lay.RC].view(B, N, lay.RC)
        out["auctioneer"] = h[:, cl.c_auctioneer: cl.c_auctioneer + self.C]
        out["agent"] = h[:, cl.c_agent: cl.c_agent + N]
        out["quality_sum"] = r[:, cl.c_quality].contiguous().view(torch.float32)
        counts = r[:, cl.c_counts]
        out["quality_cnt"] = counts & 0xFF
        out["n_accepted"] = (counts >> 8) & 0xFF
        out["n_terminated"] = (counts >> 16) & 0xFF
        out["done"] = (counts >> 24) & 0x1
        out["flags"] = (r[:, cl.c_flags] >> 25) & 0x7F   # the sticky flags share the counts word (bits 25..31)
        return out

    # ------------------------------------------------------------------ result record views
    def rewards(self, result=None):
        lay, B, N = self.layout, self.B, self.N
        r = (self.result if result is None else result)[:B]
        f = r.view(torch.float32)
        out = {}
        out["offer"] = f[:, lay.r_offer: lay.r_offer + N * lay.RL].view(B, N, lay.RL)
        out["price"] = (f[:, lay.r_price: lay.r_price + N * lay.RL].view(B, N, lay.RL)
                        if lay.r_price >= 0 else None)
        out["acceptor"] = r[:, lay.r_acceptor: lay.r_acceptor + N * lay.RC].view(B, N, lay.RC)
        out["auctioneer"] = r[:, lay.r_auctioneer: lay.r_auctioneer + self.C]
        out["agent"] = r[:, lay.r_agent: lay.r_agent + N]
        q = r[:, lay.r_quality: lay.r_quality + 2].reshape(-1).clone().view(torch.float64).view(B)
        counts = r[:, lay.r_counts]
        out["quality_sum"] = q
        out["quality_cnt"] = counts & 0xFF
        out["n_accepted"] = (counts >> 8) & 0xFF
        out["n_terminated"] = (counts >> 16) & 0xFF
        out["done"] = (counts >> 24) & 0x1
        out["flags"] = r[:, lay.r_flags]
        nw = (self.C + 1) // 2
        ai = r[:, lay.r_auctioneer_idx: lay.r_auctioneer_idx + nw].reshape(-1).clone().view(torch.int16).view(B, 2 * nw)
        out["auctioneer_idx"] = ai[:, : self.C]
        return out

    # ------------------------------------------------------------------ observations
    def _obs_buffer(self):
        if self._obs is None:
            self._obs = torch.zeros((self.layout.padded_envs, self.layout.obs_halfs), dtype=torch.int16,
                                    device=self.device)
            self._obs_ring[self._flip] = self._obs
        return self._obs

    def obs_views(self, obs=None):
        """int16 strided views into an observation record buffer (default: the env's own, as last
        written by observe() or step(..., observe=True))."""
        lay = self.layout
        B, N, Cc, Lc, NL = self.B, self.N, self.C, self.Lc, self.NL
        Wd, OH, RA, RO = 3 + 2 * NL, lay.obs_halfs, lay.o_acc_row, lay.o_off_row
        o = self._obs_buffer() if obs is None else obs
        return dict(
            acceptor=o.as_strided((B, N, Cc, Wd), (OH, Cc * RA, RA, 1), lay.o_acceptor),
            offer=o.as_strided((B, N, Lc, 2 * Cc + 2), (OH, Lc * RO, RO, 1), lay.o_offer),
            auctioneer=o.as_strided((B, Cc, Wd), (OH, RA, 1), lay.o_auctioneer))

    def observe(self, with_ids=False):
        """Dense reference-layout observations (reference src/Agent.py:148-300,
        src/Auctioneer.py:20-77) as int16 strided views into the obs record: acceptor
        [B,N,C,3+2NL], offer [B,N,L,2C+2], auctioneer [B,C,3+2NL]; with_ids adds the offer-ID
        tables ids [B,N,C,NL], auctioneer_ids [B,C,NL] (env.correspondingOfferIDs)."""
        lay = self.layout
        B, N, Cc, NL = self.B, self.N, self.C, self.NL
        obs = self._obs_buffer()
        ids = None
        if with_ids:
            if self._ids is None:
                self._ids = torch.zeros((self.B, lay.ids_halfs), dtype=torch.int16, device=self.device)
            ids = self._ids
        L.check(self.lib.msched_observe_dense(self.handle, obs.data_ptr(),
                                              None if ids is None else ids.data_ptr(), self._stream()))
        out = self.obs_views()
        if with_ids:
            out["ids"] = ids[:, : N * Cc * NL].view(B, N, Cc, NL)
            out["auctioneer_ids"] = ids[:, N * Cc * NL:].view(B, Cc, NL)
        return out

    def _cobs_buffer(self):
        if getattr(self, "_cobs", None) is None:
            self._cobs = torch.zeros((self.layout.padded_envs, self.layout.cobs_halfs), dtype=torch.int16,
                                     device=self.device)
        return self._cobs

    def compact_views(self, cobs=None):
        """int16 views into a compact observation record buffer: core [B,C,4] (ownerID, priority,
        remainingLength, jobKind; -1 = idle), slot [B,N,L,2] (priority, remainingLength; -1 = empty), offer
        [B,N,L,2] (coreID, 0 = none; offeredReward)."""
        lay = self.layout
        c = (self._cobs_buffer() if cobs is None else cobs)[: self.B]
        return dict(core=c[:, lay.c_core: lay.c_core + 4 * self.C].view(self.B, self.C, 4),
                    slot=c[:, lay.c_slot: lay.c_slot + 2 * self.NL].view(self.B, self.N, self.Lc, 2),
                    offer=c[:, lay.c_offer: lay.c_offer + 2 * self.NL].view(self.B, self.N, self.Lc, 2))

    def observe_compact(self):
        """Compact observations of the current state (msched_observe_compact): the content of the dense
        reference rows (src/Agent.py:148-300, src/Auctioneer.py:20-77) with every core, slot and pending
        offer stored once -- what the large domains (config 5) use instead of the 2.2 MB dense record."""
        L.check(self.lib.msched_observe_compact(self.handle, self._cobs_buffer().data_ptr(), self._stream()))
        return self.compact_views()

    def auctioneer_action(self, random_ties=True):
        """Auctioneer.getAuctioneerAction (reference src/Auctioneer.py:95-102) on the current
        state: int16 [B,C] table indices (N*L = reject)."""
        out = torch.empty((self.B, self.C), dtype=torch.int16, device=self.device)
        L.check(self.lib.msched_auctioneer_action(self.handle, int(random_ties), out.data_ptr(),
                                                  self._stream()))
        return out

    def hardcoded_actions(self, obs=None, random_ties=True, u=None, want_ncand=False):
        """DividedHardcodedAgent.getActions of every agent (reference src/Agent.py:622-641,
        src/HardcodedModules.py:16-45,81-109) on an observation record (default: the env's current one):
        fills the acceptor idx and offer core fields of the env's action record and returns those views
        ([B,N,C], [B,N,L]).  u: float32 [B, N*C+N*L] tie draws (parity tests); want_ncand adds the number of
        tie candidates per unit, int32 [B, N*C+N*L]."""
        obs = self._obs_buffer() if obs is None else obs
        U = self.N * self.C + self.NL
        if u is not None:
            u = torch.as_tensor(u, dtype=torch.float32).to(self.device).contiguous().view(self.B, U)
        nc = torch.zeros((self.B, U), dtype=torch.int32, device=self.device) if want_ncand else None
        L.check(self.lib.msched_hardcoded_actions(self.handle, obs.data_ptr(), int(bool(random_ties)),
                                                  None if u is None else u.data_ptr(), self.action.data_ptr(),
                                                  None if nc is None else nc.data_ptr(), self._stream()))
        self._keep_u = u
        out = (self.acceptor_actions, self.offer_core_actions)
        return out + (nc,) if want_ncand else out

    def rollout_hardcoded(self, results, obs=None, obs_every=False, random_ties=True):
        """msched_rollout_hardcoded: T = results.shape[0] x (step ; hard-coded agents on the new observations) in ONE
        launch.  The env's action record must hold the first step's actions (hardcoded_actions()); afterwards it holds
        the agents' actions for the step after the last.  results int32 [T, padded_envs, result_words]; obs: the env's
        observation record (last step) or int16 [T, padded_envs, obs_halfs] with obs_every.  Asynchronous."""
        T = int(results.shape[0])
        assert results.is_contiguous()
        if obs is None:
            assert not obs_every
            obs = self._obs_buffer()
        L.check(self.lib.msched_rollout_hardcoded(self.handle, self.action.data_ptr(), T, results.data_ptr(), obs.data_ptr(),
                                                  1 if obs_every else 0, int(bool(random_ties)), self._stream()))
        return results, obs

    # ------------------------------------------------------------------ debug / parity
    def export_state(self, env0=0, count=None):
        """Reference-shaped dump (numpy) of envs [env0, env0+count), see include/msched.h."""
        count = self.B - env0 if count is None else count
        N, Cc, Lc, NL, K = self.N, self.C, self.Lc, self.NL, self.cfg.chainCapacity
        dev = self.device
        core = torch.empty((count, Cc, 7), dtype=torch.int32, device=dev)
        slot = torch.empty((count, NL, 7), dtype=torch.int32, device=dev)
        off = torch.empty((count, NL, 5), dtype=torch.int32, device=dev)
        chain = torch.empty((count, Cc, K, 5), dtype=torch.int32, device=dev)
        clen = torch.empty((count, Cc), dtype=torch.int32, device=dev)
        misc = torch.empty((count, 4), dtype=torch.int32, device=dev)
        L.check(self.lib.msched_export_state(self.handle, env0, count, core.data_ptr(),
                                             slot.data_ptr(), off.data_ptr(), chain.data_ptr(),
                                             clen.data_ptr(), misc.data_ptr(), self._stream()))
        core, slot, off, chain, clen, misc = (t.cpu().numpy() for t in (core, slot, off, chain, clen, misc))
        s = slot.reshape(count, N, Lc, 7)
        f = off.reshape(count, N, Lc, 5)
        return dict(
            core_owner=core[..., 0], core_prio=core[..., 1], core_rem=core[..., 2],
            core_jobid=core[..., 3], core_kind=core[..., 4], core_birth=core[..., 5],
            core_init=core[..., 6],
            slot_prio=s[..., 0], slot_rem=s[..., 1], slot_jobid=s[..., 2], slot_kind=s[..., 3],
            slot_wait=s[..., 4], slot_birth=s[..., 5], slot_init=s[..., 6],
            off_core=f[..., 0], off_recip=f[..., 1], off_price=f[..., 2], off_time=f[..., 3],
            off_id=f[..., 4], chain=chain, chain_len=clen, job_counter=misc[:, 0],
            flags=misc[:, 1].astype(np.uint32))
