// coop_step_kernel.cuh -- one SchedulingEnv.step with G lanes cooperating on each environment.
//
// Same semantics and record formats as step_kernel.cuh (see the reference citations there); the
// difference is the mapping: a group of G lanes (G = 4..32, a sub-warp) owns one environment.
// Lane g of the group owns the cores j = g (mod G), the slots s = g (mod G) and the agents
// a = g (mod G); the data-parallel phases (result zeroing, per-core set-up, job progress and
// completion with the liability-chain walk, offer creation, spawn refill) run on all lanes at once,
// the offer sweep is a broadcast loop in which the owner lane of the addressed core keeps the rank /
// arg-max state (the auction's winner selection), and only executeAnOffer -- inherently serial in
// reference order -- runs on the group leader.  Groups synchronise with __syncwarp(groupMask) and
// exchange through the environment's shared-memory records; counts and flags are reduced with
// shuffles, spawn jobIDs are assigned by a ballot prefix over the agents.
//
// Small domains (config 2/3/4) use G = 4: 8 environments per warp, 4x the warps of the lane-per-env
// kernel, a much shorter critical path per environment.  Large domains (config 5: 64 cores, 256
// slots) use G = 32: one warp per environment.
#pragma once
#include "msched_common.cuh"
#include "step_kernel.cuh"

namespace msched {

// per-core scratch, 9 ints per core (record stride forced odd)
enum { SC_SEL = 0, SC_KIDX, SC_OWNER, SC_CNT, SC_BN, SC_BD, SC_NC, SC_RANK, SC_ORDER, SC_WORDS };
__host__ __device__ inline int coop_scratch_words(int C) { return make_odd(SC_WORDS * C); }

template <int G>
__device__ __forceinline__ int group_sum(int v, unsigned gmask)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_or(uint32_t v, unsigned gmask)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(gmask, v, o);
    return v;
}

template <int G>
__device__ __forceinline__ void coop_step_env(const DevParams &p, uint32_t *st, const int16_t *act,
                                              uint32_t *res, int *scr, int env, int g, unsigned gmask)
{
    const int N = p.N, C = p.C, L = p.L, NL = p.NL;
    const int mode = p.mode;
    const bool agg = mode == MSCHED_REWARD_AGGREGATED_FIXED;
    const bool freeM = mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL ||
                       mode == MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL;
    const bool external = p.auctionMode == MSCHED_AUCTION_EXTERNAL;
    const int round = cur_round(p);
    uint32_t *core = st + 2;
    unsigned char *chl = reinterpret_cast<unsigned char *>(st + p.sChlen);
    uint32_t *slot = st + p.sSlot;
    float *resf = reinterpret_cast<float *>(res);
    int *resi = reinterpret_cast<int *>(res);
    uint32_t flags = 0u;

    // ---- phase 0: zero the result record, range-check the acceptor actions, chain prefetch ----
    for (int k = g; k < p.RW; k += G) res[k] = 0u;
    for (int k = g; k < N * C; k += G) {
        const int a = act[p.aAcc + k];
        if (a < 0 || a > NL) flags |= MSCHED_FLAG_ACTION_RANGE;
    }
    // ---- phase 1 (pass A): who acts on each core and with which table index ----
    for (int j = g; j < C; j += G) {
        const uint32_t cw0 = core[3 * j];
        const int o = core_owner(cw0);
        int k = -1;
        if (o > 0) k = act[p.aAcc + (o - 1) * C + j];
        else if (external) {
            k = act[p.aAuc + j];
            if (k < 0 || k > NL) flags |= MSCHED_FLAG_ACTION_RANGE;
        }
        int *s = scr + SC_WORDS * j;
        s[SC_SEL] = -1;
        s[SC_KIDX] = (k >= 0 && k < NL) ? k : -1;
        s[SC_OWNER] = o;
        s[SC_CNT] = 0;
        s[SC_BN] = -1;
        s[SC_BD] = 1;
        s[SC_NC] = 0;
        s[SC_RANK] = 0;
        if (job_kind(cw0) >= 0 && job_rem(cw0) == 1)
            prefetch_l1(reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap);
    }
    __syncwarp(gmask);

    // ---- phase 2 (pass B): sweep the pending offers in creation order; the owner lane of the
    // addressed core ranks them and keeps the auction's running arg-max ----
    for (int s = 0; s < NL; ++s) {
        const uint32_t w3 = slot[4 * s + 3];  // same address for the whole group: broadcast
        const int c = (int)(w3 & 0xffu);
        if (c == 0) continue;
        const int j = c - 1;
        if ((j & (G - 1)) != g) continue;
        int *sc = scr + SC_WORDS * j;
        const int r = off_recip(w3);
        if (r != sc[SC_OWNER]) continue;
        const int rank = sc[SC_CNT];
        sc[SC_CNT] = rank + 1;
        if (r > 0 || external) {
            if (rank == sc[SC_KIDX]) { sc[SC_SEL] = s; sc[SC_RANK] = rank; }
        } else {
            int pn = off_price(w3), pd = job_rem(slot[4 * s]);
            if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
            const int bn = sc[SC_BN], bd = sc[SC_BD];
            const int lhs = pn * bd, rhs = bn * pd;
            if (lhs > rhs) {
                sc[SC_BN] = pn; sc[SC_BD] = pd; sc[SC_NC] = 1; sc[SC_SEL] = s; sc[SC_RANK] = rank;
            } else if (lhs == rhs) {
                sc[SC_NC] += 1;
            }
        }
    }
    // auction epilogue (random arg-max) + auctioneer index report + execution key, per owned core
    for (int j = g; j < C; j += G) {
        int *sc = scr + SC_WORDS * j;
        int kUsed = NL;
        if (external) {
            kUsed = act[p.aAuc + j];
        } else if (sc[SC_OWNER] == 0 && sc[SC_SEL] >= 0) {
            const int ncand = sc[SC_NC];
            if (p.auctionMode == MSCHED_AUCTION_RANDOM_MAX && ncand > 1) {
                uint32_t x[4];
                env_draw(p, env, kStreamTie, (uint32_t)(j >> 2), 0u, x);  // word j%4 of call j/4
                const uint32_t xw = (j & 3) == 0 ? x[0] : (j & 3) == 1 ? x[1] : (j & 3) == 2 ? x[2] : x[3];
                int pick = (int)__umulhi(xw, (uint32_t)ncand);
                if (pick > 0) {
                    const int bn = sc[SC_BN], bd = sc[SC_BD];
                    int rank = 0;
                    for (int s = 0; s < NL; ++s) {
                        const uint32_t w3 = slot[4 * s + 3];
                        if ((w3 & 0xffffu) != (uint32_t)(j + 1)) continue;
                        int pn = off_price(w3), pd = job_rem(slot[4 * s]);
                        if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                        if (pn * bd == bn * pd) {
                            if (pick == 0) { sc[SC_SEL] = s; sc[SC_RANK] = rank; break; }
                            --pick;
                        }
                        ++rank;
                    }
                }
            }
            kUsed = sc[SC_RANK];
        }
        // 16-bit halves of one word may belong to two lanes: atomic OR
        atomicOr(&res[p.rAucIdx + (j >> 1)], ((uint32_t)(kUsed & 0xffff)) << ((j & 1) * 16));
    }
    __syncwarp(gmask);

    // ---- execution order: rank of (owner, core) among the cores with a selected offer ----
    int nExecLocal = 0;
    for (int j = g; j < C; j += G) {
        const int *sc = scr + SC_WORDS * j;
        if (sc[SC_SEL] < 0) continue;
        const int o = sc[SC_OWNER];
        const int key = (((o == 0) ? (N + 1) : o) << 8) | j;
        int rank = 0;
        for (int i = 0; i < C; ++i) {
            const int *si = scr + SC_WORDS * i;
            if (si[SC_SEL] < 0) continue;
            const int oi = si[SC_OWNER];
            rank += ((((oi == 0) ? (N + 1) : oi) << 8) | i) < key;
        }
        scr[SC_WORDS * rank + SC_ORDER] = j;
        ++nExecLocal;
    }
    const int nExec = group_sum<G>(nExecLocal, gmask);
    __syncwarp(gmask);

    // ---- phase 3: executeAnOffer, serial in reference order, on the group leader ----
    double qualSum = 0.0;
    int qualCnt = 0;
    if (g == 0) {
        for (int e = 0; e < nExec; ++e) {
            const int j = scr[SC_WORDS * e + SC_ORDER];
            const int sel = scr[SC_WORDS * j + SC_SEL];
            const int who = scr[SC_WORDS * j + SC_OWNER];
            const int selA = sel / L;
            const uint32_t cw0 = core[3 * j], cw1 = core[3 * j + 1], cw2 = core[3 * j + 2];
            const uint32_t sw0 = slot[4 * sel], sw1 = slot[4 * sel + 1], sw2 = slot[4 * sel + 2];
            const uint32_t sw3 = slot[4 * sel + 3];
            const int kind = job_kind(sw0), time = job_rem(sw0), price = off_price(sw3);
            const int offerer = selA + 1;
            const int prio1 = p.prio[kind];
            stat_accept(p, env, kind, price);
            slot[4 * sel] = kEmptyJobW0; slot[4 * sel + 1] = kEmptyId; slot[4 * sel + 2] = kEmptyId;
            slot[4 * sel + 3] = 0u;
            core[3 * j] = pack_core(offerer, kind, time);
            core[3 * j + 1] = sw1;
            core[3 * j + 2] = sw2;
            if (who > 0) {
                const int base = (who - 1) * L;
                int q = -1;
                for (int t = 0; t < L; ++t)
                    if (job_kind(slot[4 * (base + t)]) < 0) { q = t; break; }
                if (q >= 0) {
                    slot[4 * (base + q)] = cw0 & 0xffffff00u;
                    slot[4 * (base + q) + 1] = cw1;
                    slot[4 * (base + q) + 2] = cw2;
                    slot[4 * (base + q) + 3] = 0u;
                } else {
                    flags |= MSCHED_FLAG_COLLECTION_FULL;
                }
                double qv = __dmul_rn((double)price, c_rcp[time]);
                const int fk = job_kind(cw0);
                if (fk >= 0) qv = __dsub_rn(qv, __dmul_rn((double)p.prio[fk], c_rcp[job_rem(cw0) & 0xff]));
                qualSum = __dadd_rn(qualSum, __dmul_rn(qv, 10.0));
                ++qualCnt;
            }
            const int len = chl[j];
            if (len < p.chainCap) {
                uint2 *ce = reinterpret_cast<uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap + len;
                *ce = make_uint2((uint32_t)round, pack_chain(price, time, offerer));
                chl[j] = (unsigned char)(len + 1);
            } else {
                flags |= MSCHED_FLAG_CHAIN_OVERFLOW;
            }
            if (agg) {
                resf[p.rOffer + selA] += (float)prio1;
            } else {
                resf[p.rOffer + sel] = (float)prio1;
                if (freeM) {
                    const int df = prio1 - price;
                    float pr;
                    if (mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL)
                        pr = (df == 0) ? p.netZero : (float)df;
                    else
                        pr = (df >= 0) ? (float)prio1 : (float)df;
                    resf[p.rPrice + sel] = pr;
                }
            }
        }
    }
    __syncwarp(gmask);

    // ---- phase 4: job progress / completion + termination rewards, one lane per core ----
    int nTerm = 0;
    for (int j = g; j < C; j += G) {
        const uint32_t cw0 = core[3 * j];
        const int kind = job_kind(cw0);
        if (kind < 0) continue;
        const int rem = job_rem(cw0) - 1;
        if (rem != 0) {
            core[3 * j] = (cw0 & 0x0000ffffu) | ((uint32_t)rem << 16);
            continue;
        }
        const int R = p.mult * p.prio[kind];
        const int o = core_owner(cw0) - 1;
        stat_terminate(p, env, kind, round, core[3 * j + 2]);
        if (agg) {
            atomicAdd(&resi[p.rAcc + o], R);
            atomicAdd(&resi[p.rAgent + o], R);
        } else {
            resi[p.rAcc + o * C + j] = R;  // column j belongs to this lane
            if (!freeM) atomicAdd(&resi[p.rAgent + o], R);
        }
        const int len = chl[j];
        const uint2 *ce = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
        int recip = 0;
        for (int e = 0; e < len; ++e) {
            const uint2 en = ce[e];
            const int price = (int)(int16_t)(en.y & 0xffffu);
            const int time = (int)((en.y >> 16) & 0xffu);
            const int offerer = (int)(en.y >> 24);
            const int traded = traded_reward(price, time, (round + 1) - (int)en.x);
            atomicAdd(&resi[p.rAgent + offerer - 1], -traded);
            if (agg) {
                atomicAdd(&resi[p.rAcc + offerer - 1], -traded);
                if (recip > 0) atomicAdd(&resi[p.rAgent + recip - 1], traded);
            } else {
                resi[p.rAcc + (offerer - 1) * C + j] -= traded;
                if (recip > 0) {
                    resi[p.rAcc + (recip - 1) * C + j] += traded;
                    atomicAdd(&resi[p.rAgent + recip - 1], traded);
                }
            }
            if (recip == 0) resi[p.rAuc + j] = traded;
            recip = offerer;
        }
        chl[j] = 0;
        core[3 * j] = kEmptyJobW0;
        core[3 * j + 1] = kEmptyId;
        core[3 * j + 2] = kEmptyId;
        ++nTerm;
    }
    __syncwarp(gmask);

    // ---- phase 5: offer creation, one lane per slot ----
    for (int s = g; s < NL; s += G) {
        const int a = act[p.aOffc + s];
        const uint32_t w0 = slot[4 * s];
        const int kind = job_kind(w0);
        const bool waitOld = (slot[4 * s + 3] & 0xffu) != 0u;
        uint32_t w3 = 0u;
        if (a >= 0 && a < C && kind >= 0 && !waitOld) {
            const int price = p.freePrices ? (int)act[p.aOffp + s] : p.fix[kind];
            w3 = pack_offer(a + 1, core_owner(core[3 * a]), price);
        }
        slot[4 * s + 3] = w3;
    }
    __syncwarp(gmask);

    // ---- phase 6: spawn refill, one lane per agent; jobIDs by ballot prefix in agent order ----
    uint32_t jobctr = st[0];
    const unsigned laneInGroupLt = (1u << g) - 1u;
    const int shift = (threadIdx.x & 31) & ~(G - 1);
    for (int a0 = 0; a0 < N; a0 += G) {
        const int a = a0 + g;
        bool spawn = false;
        if (a < N) {
            int owned = 0, nfree = 0;
            for (int j = 0; j < C; ++j) owned += (core_owner(core[3 * j]) == a + 1);
            for (int q = 0; q < L; ++q) nfree += (job_kind(slot[4 * (a * L + q)]) < 0);
            spawn = owned + p.newJobs <= nfree;
        }
        const unsigned bal = (__ballot_sync(gmask, spawn) >> shift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
        if (spawn) {
            uint32_t id = jobctr + (uint32_t)(__popc(bal & laneInGroupLt) * p.newJobs);
            uint32_t rnd[4] = {0u, 0u, 0u, 0u};
            int rndCall = -1;
            for (int k = 0; k < p.newJobs; ++k) {
                int kind = -1;
                if (p.spawnMode == MSCHED_SPAWN_KINDS) {
                    kind = act[p.aSpawn + a * p.newJobs + k];
                } else {
                    double u;
                    if (p.spawnMode == MSCHED_SPAWN_U64) {
                        u = p.spawnU[((size_t)env * N + a) * p.newJobs + k];
                    } else {
                        const int dnum = a * p.newJobs + k;
                        if ((dnum >> 2) != rndCall) {
                            rndCall = dnum >> 2;
                            env_draw(p, env, kStreamSpawn, (uint32_t)rndCall, 0u, rnd);
                        }
                        const uint32_t xr = (dnum & 3) == 0 ? rnd[0] : (dnum & 3) == 1 ? rnd[1]
                                          : (dnum & 3) == 2 ? rnd[2] : rnd[3];
                        u = (double)xr * (1.0 / 4294967296.0);
                    }
                    for (int q = 0; q < p.J; ++q)
                        if (u < p.cum[q]) { kind = q; break; }
                }
                if (kind < 0 || kind >= p.J) { flags |= MSCHED_FLAG_SPAWN_RANGE; kind = p.J - 1; }
                int q = 0;
                for (int t = L - 1; t >= 0; --t)
                    if (job_kind(slot[4 * (a * L + t)]) < 0) q = t;
                const int s = a * L + q;
                slot[4 * s] = pack_slot(kind, p.len[kind]);
                slot[4 * s + 1] = id++;
                slot[4 * s + 2] = (uint32_t)round;
                slot[4 * s + 3] = 0u;
            }
        }
        jobctr += (uint32_t)(__popc(bal) * p.newJobs);
    }

    // ---- phase 7: scalar outputs ----
    flags = group_or<G>(flags, gmask);
    nTerm = group_sum<G>(nTerm, gmask);
    if (g == 0) {
        flags |= st[1];
        st[0] = jobctr;
        st[1] = flags;
        const unsigned long long qb = (unsigned long long)__double_as_longlong(qualSum);
        res[p.rQual] = (uint32_t)qb;
        res[p.rQual + 1] = (uint32_t)(qb >> 32);
        res[p.rCounts] = (uint32_t)qualCnt | ((uint32_t)nExec << 8) | ((uint32_t)nTerm << 16) |
                         ((uint32_t)cur_done(p, round) << 24);
        res[p.rFlags] = flags;
    }
}

// grid = Bpad / (blockDim.x / G) tiles
template <int G>
__global__ void __launch_bounds__(128) coop_step_kernel(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int E = blockDim.x / G;  // envs per CTA
    const int tid = threadIdx.x;
    const int env0 = blockIdx.x * E;
    const uint32_t stBytes = (uint32_t)E * p.W * 4u, acBytes = (uint32_t)E * p.AH * 2u,
                   rsBytes = (uint32_t)E * p.RW * 4u;
    // the result tile must stay 16-byte aligned: pad the action tile
    const uint32_t acPad = (acBytes + 15u) & ~15u;
    uint32_t *sState = reinterpret_cast<uint32_t *>(smem);
    int16_t *sAct = reinterpret_cast<int16_t *>(smem + stBytes);
    uint32_t *sRes = reinterpret_cast<uint32_t *>(smem + stBytes + acPad);
    const int SCR = coop_scratch_words(p.C);
    int *sScr = reinterpret_cast<int *>(smem + stBytes + acPad + rsBytes);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, stBytes + acBytes);
        bulk_g2s(sState, p.state + (size_t)env0 * p.W, stBytes, &bar);
        bulk_g2s(sAct, p.action + (size_t)env0 * p.AH, acBytes, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    const int e = tid / G, g = tid % G;
    const int lane = tid & 31;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int env = env0 + e;
    if (env < p.B) {
        coop_step_env<G>(p, sState + (size_t)e * p.W, sAct + (size_t)e * p.AH, sRes + (size_t)e * p.RW,
                         sScr + (size_t)e * SCR, env, g, gmask);
    } else {
        for (int k = g; k < p.RW; k += G) sRes[(size_t)e * p.RW + k] = 0u;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(p.state + (size_t)env0 * p.W, sState, stBytes);
        bulk_s2g(p.result + (size_t)env0 * p.RW, sRes, rsBytes);
        bulk_commit();
        bulk_wait_read();
        finish_round(p);
    }
}

}  // namespace msched
