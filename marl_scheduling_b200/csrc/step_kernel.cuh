// step_kernel.cuh -- one SchedulingEnv.step for a tile of environments, one lane per env.
//
// Reference semantics restated (file:line relative to the reference):
//   src/SchedulingEnvironment.py:32-83   SchedulingEnv.step
//   src/world.py:295-334                 World.step1 (order of phases)
//   src/world.py:391-404, 378-389        executeAgentAcceptions1 / executeAuctioneerAcceptions
//   src/world.py:261-293                 executeAnOffer
//   src/HardcodedModules.py:48-78        HardcodedAuctioneerAcceptor.selectAction (the auction)
//   src/world.py:336-367                 processOneTimestepAndUpdateOwnership
//   src/world.py:406-478                 create{Fix,Free}PriceOfferObjectsFromActions
//   src/world.py:369-376, src/Agent.py:50-70  spawn refill
//   src/Reward.py:6-212                  the three reward functions
//   src/SchedulingEnvironment.py:174-192 acception quality
//
// Mapping: a CTA owns a tile of blockDim.x consecutive environments.  The tile's state records
// and action records are each ONE contiguous HBM chunk; thread 0 moves them into shared memory
// with cp.async.bulk (TMA) and every lane then advances its own environment entirely out of
// shared memory (record strides are odd, so any field index is bank-conflict free).  Results
// are built in a shared-memory result tile and written back with two bulk stores.  The only
// scattered HBM traffic is the liability chain: one 8-byte append per acceptance, one short
// walk per job completion.
#pragma once
#include "msched_common.cuh"
#include "reward_math.cuh"

namespace msched {

template <int TN, int TC, int TL>
struct Dims {
    const int N, C, L, NL;
    __device__ __forceinline__ explicit Dims(const DevParams &p)
        : N(TN ? TN : p.N), C(TC ? TC : p.C), L(TL ? TL : p.L), NL((TN ? TN : p.N) * (TL ? TL : p.L))
    {
    }
};

// per-lane scratch (shared memory, 3 words per core, record stride odd):
//   w0 = selected slot (0xff none) | offers seen << 8 | acceptor index (0xff none) << 16 | owner << 24
//   w1 = best price i16 | best time u8 << 16 | #candidates << 24      (in-kernel auction)
//   w2 = rank of the selected offer inside the core's offer table
__host__ __device__ inline int scratch_words(int C) { return make_odd(3 * C); }

template <int TN, int TC, int TL>
__device__ __forceinline__ void step_env(const DevParams &p, uint32_t *__restrict__ st,
                                         const int16_t *__restrict__ act, uint32_t *__restrict__ res,
                                         uint32_t *__restrict__ scr, int env)
{
    const Dims<TN, TC, TL> d(p);
    const int N = d.N, C = d.C, L = d.L, NL = d.NL;
    const int mode = p.mode;
    const bool agg = mode == MSCHED_REWARD_AGGREGATED_FIXED;
    const bool freeM = mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL ||
                       mode == MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL;
    const bool external = p.auctionMode == MSCHED_AUCTION_EXTERNAL;
    const int round = cur_round(p);
    uint32_t *core = st + 2;
    uint32_t *chl = st + p.sChlen;
    uint32_t *slot = st + p.sSlot;
    float *resf = reinterpret_cast<float *>(res);
    int *resi = reinterpret_cast<int *>(res);
    uint32_t flags = st[1];

    // liability chains of cores whose job completes this step are walked further down; their
    // loads are cold, dependent HBM accesses, so issue them NOW: with a compile-time domain the
    // first two entries (16 B) of every candidate core go straight into registers, otherwise the
    // line is prefetched
    constexpr int PC = TC > 0 ? TC : 1;
    uint4 pre[PC];
    bool preOk[PC];
#pragma unroll
    for (int j = 0; j < PC; ++j) { pre[j] = make_uint4(0u, 0u, 0u, 0u); preOk[j] = false; }
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const uint32_t cw0 = core[3 * j];
        if (job_kind(cw0) >= 0 && job_rem(cw0) == 1) {
            const uint2 *cb = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
            if (TC > 0 && (p.chainCap & 1) == 0) {
                pre[TC > 0 ? j : 0] = *reinterpret_cast<const uint4 *>(cb);
                preOk[TC > 0 ? j : 0] = true;
            } else {
                prefetch_l1(cb);
            }
        }
    }

#pragma unroll 4
    for (int k = 0; k < p.RW; ++k) res[k] = 0u;

    // ---- range check of every acceptor action (assert in src/world.py:389,404) ----
    {
        bool bad = false;
#pragma unroll
        for (int k = 0; k < N * C; ++k) {
            const int a = act[p.aAcc + k];
            bad |= (a < 0) | (a > NL);
        }
        if (external) {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int a = act[p.aAuc + j];
                bad |= (a < 0) | (a > NL);
            }
        }
        if (bad) flags |= MSCHED_FLAG_ACTION_RANGE;
    }

    // ---- pass A: per core, who acts on it and with which table index ----
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const int o = core_owner(core[3 * j]);
        int k = -1;
        if (o > 0) k = act[p.aAcc + (o - 1) * C + j];
        else if (external) k = act[p.aAuc + j];
        const uint32_t kk = (k >= 0 && k < NL) ? (uint32_t)k : 0xffu;
        scr[3 * j] = 0xffu | (kk << 16) | ((uint32_t)o << 24);
        scr[3 * j + 1] = 0x0001ffffu;  // best = -1/1, no candidates
        scr[3 * j + 2] = 0u;
    }

    // ---- pass B: one sweep over the pending offers in creation order (agent asc, slot asc):
    // rank inside the (recipient, core) table -> selection by the acceptor index, or, for idle
    // cores in in-kernel auction mode, the running arg-max of offeredReward/necessaryTime compared
    // exactly by cross-multiplication; -1/-2 operands rate -1 (HardcodedModules.calculateRewardRatio)
#pragma unroll
    for (int s = 0; s < NL; ++s) {
        const uint32_t w3 = slot[4 * s + 3];
        const int c = (int)(w3 & 0xffu);
        if (c == 0) continue;
        const int j = c - 1;
        const int r = off_recip(w3);
        uint32_t x0 = scr[3 * j];
        if (r != (int)(x0 >> 24)) continue;  // not addressed to the core's owner: in nobody's table
        const uint32_t rank = (x0 >> 8) & 0xffu;
        x0 += 0x100u;
        if (r > 0 || external) {
            if (rank == ((x0 >> 16) & 0xffu)) { x0 = (x0 & ~0xffu) | (uint32_t)s; scr[3 * j + 2] = rank; }
        } else {
            int pn = off_price(w3), pd = job_rem(slot[4 * s]);
            if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
            uint32_t x1 = scr[3 * j + 1];
            const int bn = (int)(int16_t)(x1 & 0xffffu), bd = (int)((x1 >> 16) & 0xffu);
            const int lhs = pn * bd, rhs = bn * pd;
            if (lhs > rhs) {
                x1 = (uint32_t)(pn & 0xffff) | ((uint32_t)pd << 16) | (1u << 24);
                x0 = (x0 & ~0xffu) | (uint32_t)s;
                scr[3 * j + 2] = rank;
            } else if (lhs == rhs) {
                x1 += 1u << 24;
            }
            scr[3 * j + 1] = x1;
        }
        scr[3 * j] = x0;
    }

    // ---- auction epilogue: uniformly random arg-max (random.sample in the reference), and the
    // auctioneer index per core as Auctioneer.getAuctioneerAction reports it ----
#pragma unroll
    for (int j = 0; j < C; ++j) {
        uint32_t x0 = scr[3 * j];
        int kUsed = NL;
        if (external) {
            kUsed = act[p.aAuc + j];
        } else if ((x0 >> 24) == 0u && (x0 & 0xffu) != 0xffu) {
            const uint32_t x1 = scr[3 * j + 1];
            const int ncand = (int)(x1 >> 24);
            if (p.auctionMode == MSCHED_AUCTION_RANDOM_MAX && ncand > 1) {
                uint32_t x[4];
                env_draw(p, env, kStreamTie, (uint32_t)(j >> 2), 0u, x);  // word j%4 of call j/4
                const uint32_t xw = (j & 3) == 0 ? x[0] : (j & 3) == 1 ? x[1] : (j & 3) == 2 ? x[2] : x[3];
                int pick = (int)__umulhi(xw, (uint32_t)ncand);
                if (pick > 0) {
                    const int bn = (int)(int16_t)(x1 & 0xffffu), bd = (int)((x1 >> 16) & 0xffu);
                    int rank = 0;
                    for (int s = 0; s < NL; ++s) {
                        const uint32_t w3 = slot[4 * s + 3];
                        if ((w3 & 0xffffu) != (uint32_t)(j + 1)) continue;
                        int pn = off_price(w3), pd = job_rem(slot[4 * s]);
                        if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                        if (pn * bd == bn * pd) {
                            if (pick == 0) {
                                x0 = (x0 & ~0xffu) | (uint32_t)s;
                                scr[3 * j] = x0;
                                scr[3 * j + 2] = (uint32_t)rank;
                                break;
                            }
                            --pick;
                        }
                        ++rank;
                    }
                }
            }
            kUsed = (int)scr[3 * j + 2];
        }
        res[p.rAucIdx + (j >> 1)] |= ((uint32_t)(kUsed & 0xffff)) << ((j & 1) * 16);
    }

    double qualSum = 0.0;
    int qualCnt = 0, nAcc = 0, nTerm = 0;

    // ---- executeAnOffer in reference order: agents ascending, cores ascending, auctioneer last
    // (at most one acceptance per core, invariant I2) ----
    int lastKey = -1;
    for (int e = 0; e < C; ++e) {
        int best = 0x7fffffff, j = -1;
#pragma unroll
        for (int jj = 0; jj < C; ++jj) {
            const uint32_t x0 = scr[3 * jj];
            const int o = (int)(x0 >> 24);
            const int key = (((o == 0) ? (N + 1) : o) << 8) | jj;
            if ((x0 & 0xffu) != 0xffu && key > lastKey && key < best) { best = key; j = jj; }
        }
        if (j < 0) break;
        lastKey = best;
        const int sel = (int)(scr[3 * j] & 0xffu);
        const int who = (int)(scr[3 * j] >> 24);
        const int selA = sel / L, selQ = sel - selA * L;
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
            // old job back into the recipient's first empty slot
            const int base = (who - 1) * L;
            int q = -1;
#pragma unroll
            for (int t = L - 1; t >= 0; --t)
                if (job_kind(slot[4 * (base + t)]) < 0) q = t;
            if (q >= 0) {
                slot[4 * (base + q)] = cw0 & 0xffffff00u;
                slot[4 * (base + q) + 1] = cw1;
                slot[4 * (base + q) + 2] = cw2;
                slot[4 * (base + q) + 3] = 0u;
            } else {
                flags |= MSCHED_FLAG_COLLECTION_FULL;
            }
            // acception quality, src/SchedulingEnvironment.py:174-192 (former = core before)
            double qv = __dmul_rn((double)price, c_rcp[time]);
            const int fk = job_kind(cw0);
            if (fk >= 0) qv = __dsub_rn(qv, __dmul_rn((double)p.prio[fk], c_rcp[job_rem(cw0) & 0xff]));
            qualSum = __dadd_rn(qualSum, __dmul_rn(qv, 10.0));
            ++qualCnt;
        }
        {  // liability chain append (stored oldest first)
            const uint32_t cw = chl[j >> 2];
            const int len = (int)((cw >> ((j & 3) * 8)) & 0xffu);
            if (len < p.chainCap) {
                uint2 *ce = reinterpret_cast<uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap + len;
                *ce = make_uint2((uint32_t)round, pack_chain(price, time, offerer));
                chl[j >> 2] = cw + (1u << ((j & 3) * 8));
#pragma unroll
                for (int jj = 0; jj < PC; ++jj)
                    if (jj == j) preOk[jj] = false;  // the preloaded copy is stale now
            } else {
                flags |= MSCHED_FLAG_CHAIN_OVERFLOW;
            }
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
        (void)selQ;
        ++nAcc;
    }

    // ---- job progress / completion + termination rewards ----
#pragma unroll
    for (int j = 0; j < C; ++j) {
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
            resi[p.rAcc + o] += R;
            resi[p.rAgent + o] += R;
        } else {
            resi[p.rAcc + o * C + j] = R;
            if (!freeM) resi[p.rAgent + o] += R;
        }
        const uint32_t cw = chl[j >> 2];
        const int len = (int)((cw >> ((j & 3) * 8)) & 0xffu);
        const uint2 *ce = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
        int recip = 0;  // the oldest entry was accepted by the auctioneer
        for (int e = 0; e < len; ++e) {
            uint2 en;
            if (TC > 0 && e < 2 && preOk[TC > 0 ? j : 0]) {
                const uint4 q4 = pre[TC > 0 ? j : 0];
                en = e == 0 ? make_uint2(q4.x, q4.y) : make_uint2(q4.z, q4.w);
            } else {
                en = ce[e];
            }
            const int price = (int)(int16_t)(en.y & 0xffffu);
            const int time = (int)((en.y >> 16) & 0xffu);
            const int offerer = (int)(en.y >> 24);
            const int traded = traded_reward(price, time, (round + 1) - (int)en.x);
            resi[p.rAgent + offerer - 1] -= traded;
            if (agg) {
                resi[p.rAcc + offerer - 1] -= traded;
                if (recip > 0) resi[p.rAgent + recip - 1] += traded;
            } else {
                resi[p.rAcc + (offerer - 1) * C + j] -= traded;
                if (recip > 0) {
                    resi[p.rAcc + (recip - 1) * C + j] += traded;
                    resi[p.rAgent + recip - 1] += traded;
                }
            }
            if (recip == 0) resi[p.rAuc + j] = traded;
            recip = offerer;
        }
        chl[j >> 2] = cw & ~(0xffu << ((j & 3) * 8));
        core[3 * j] = kEmptyJobW0;
        core[3 * j + 1] = kEmptyId;
        core[3 * j + 2] = kEmptyId;
        ++nTerm;
    }

    // ---- offer creation ----
#pragma unroll
    for (int s = 0; s < NL; ++s) {
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

    // ---- spawn refill ----
    uint32_t jobctr = st[0];
    uint32_t rnd[4] = {0u, 0u, 0u, 0u};
    int rndCall = -1;
#pragma unroll
    for (int a = 0; a < N; ++a) {
        int owned = 0, nfree = 0;
#pragma unroll
        for (int j = 0; j < C; ++j) owned += (core_owner(core[3 * j]) == a + 1);
#pragma unroll
        for (int q = 0; q < L; ++q) nfree += (job_kind(slot[4 * (a * L + q)]) < 0);
        if (owned + p.newJobs > nfree) continue;
        for (int k = 0; k < p.newJobs; ++k) {
            int kind = -1;
            if (p.spawnMode == MSCHED_SPAWN_KINDS) {
                kind = act[p.aSpawn + a * p.newJobs + k];
            } else {
                double u;
                if (p.spawnMode == MSCHED_SPAWN_U64) {
                    u = p.spawnU[((size_t)env * N + a) * p.newJobs + k];
                } else {
                    const int dnum = a * p.newJobs + k;  // draw d uses word d%4 of Philox call d/4
                    if ((dnum >> 2) != rndCall) {
                        rndCall = dnum >> 2;
                        env_draw(p, env, kStreamSpawn, (uint32_t)rndCall, 0u, rnd);
                    }
                    const uint32_t xr = (dnum & 3) == 0 ? rnd[0] : (dnum & 3) == 1 ? rnd[1] : (dnum & 3) == 2 ? rnd[2] : rnd[3];
                    u = (double)xr * (1.0 / 4294967296.0);
                }
                for (int q = 0; q < p.J; ++q)
                    if (u < p.cum[q]) { kind = q; break; }
            }
            if (kind < 0 || kind >= p.J) { flags |= MSCHED_FLAG_SPAWN_RANGE; kind = p.J - 1; }
            int q = 0;
#pragma unroll
            for (int t = L - 1; t >= 0; --t)
                if (job_kind(slot[4 * (a * L + t)]) < 0) q = t;
            const int s = a * L + q;  // an empty slot exists by the guard above
            slot[4 * s] = pack_slot(kind, p.len[kind]);
            slot[4 * s + 1] = jobctr++;
            slot[4 * s + 2] = (uint32_t)round;
            slot[4 * s + 3] = 0u;
        }
    }
    st[0] = jobctr;
    st[1] = flags;

    // ---- scalar outputs ----
    const unsigned long long qb = (unsigned long long)__double_as_longlong(qualSum);
    res[p.rQual] = (uint32_t)qb;
    res[p.rQual + 1] = (uint32_t)(qb >> 32);
    res[p.rCounts] = (uint32_t)qualCnt | ((uint32_t)nAcc << 8) | ((uint32_t)nTerm << 16) | ((uint32_t)cur_done(p, round) << 24);
    res[p.rFlags] = flags;
}

// grid = Bpad / blockDim.x tiles; dynamic smem = blockDim.x * (W*4 + AH*2 + RW*4 + SCR*4) bytes
template <int TN, int TC, int TL>
__global__ void __launch_bounds__(128) step_kernel(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int T = blockDim.x;
    const int lane = threadIdx.x;
    const int env0 = blockIdx.x * T;
    const uint32_t stBytes = (uint32_t)T * p.W * 4u, acBytes = (uint32_t)T * p.AH * 2u,
                   rsBytes = (uint32_t)T * p.RW * 4u;
    uint32_t *sState = reinterpret_cast<uint32_t *>(smem);
    int16_t *sAct = reinterpret_cast<int16_t *>(smem + stBytes);
    uint32_t *sRes = reinterpret_cast<uint32_t *>(smem + stBytes + acBytes);
    const int SCR = scratch_words(p.C);
    uint32_t *sScr = reinterpret_cast<uint32_t *>(smem + stBytes + acBytes + rsBytes);

    if (lane == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, stBytes + acBytes);
        bulk_g2s(sState, p.state + (size_t)env0 * p.W, stBytes, &bar);
        bulk_g2s(sAct, p.action + (size_t)env0 * p.AH, acBytes, &bar);
    }
    __syncthreads();  // barrier init visible before anyone polls it
    mbar_wait(&bar, 0);

    const int env = env0 + lane;
    if (env < p.B)
        step_env<TN, TC, TL>(p, sState + (size_t)lane * p.W, sAct + (size_t)lane * p.AH,
                             sRes + (size_t)lane * p.RW, sScr + (size_t)lane * SCR, env);
    else
        for (int k = 0; k < p.RW; ++k) sRes[(size_t)lane * p.RW + k] = 0u;

    fence_async_smem();
    __syncthreads();
    if (lane == 0) {
        bulk_s2g(p.state + (size_t)env0 * p.W, sState, stBytes);
        bulk_s2g(p.result + (size_t)env0 * p.RW, sRes, rsBytes);
        bulk_commit();
        bulk_wait_read();
        finish_round(p);
    }
}

__global__ void bump_round_kernel(int *roundDev) { *roundDev += 1; }

// World.__init__ (src/world.py:210-254): every core idle and auctioneer-owned, every slot empty
__global__ void reset_kernel(const __grid_constant__ DevParams p)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.Bpad) return;
    uint32_t *st = p.state + (size_t)env * p.W;
    for (int k = 0; k < p.W; ++k) st[k] = 0u;
    st[0] = 1u;
    for (int j = 0; j < p.C; ++j) {
        st[2 + 3 * j] = kEmptyJobW0;
        st[2 + 3 * j + 1] = kEmptyId;
        st[2 + 3 * j + 2] = kEmptyId;
    }
    for (int s = 0; s < p.NL; ++s) {
        st[p.sSlot + 4 * s] = kEmptyJobW0;
        st[p.sSlot + 4 * s + 1] = kEmptyId;
        st[p.sSlot + 4 * s + 2] = kEmptyId;
        st[p.sSlot + 4 * s + 3] = 0u;
    }
}

}  // namespace msched
