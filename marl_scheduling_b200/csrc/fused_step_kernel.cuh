// fused_step_kernel.cuh -- SchedulingEnv.step AND the dense observations of the new state in one
// launch, for compile-time domains (N agents, C cores, L slots per agent).
//
// Same reference semantics as step_kernel.cuh:
//   src/SchedulingEnvironment.py:32-83   SchedulingEnv.step
//   src/world.py:295-334                 World.step1 (order of phases)
//   src/world.py:261-293, 378-404        executeAnOffer, agent / auctioneer acceptances
//   src/HardcodedModules.py:48-78        the auction rule
//   src/world.py:336-367, 406-478        progress/completion, offer creation
//   src/world.py:369-376, src/Agent.py:50-70   spawn refill
//   src/Reward.py:6-212                  rewards
//   src/Agent.py:148-300, src/Auctioneer.py:20-77   observations of the state AFTER the step
//
// Mapping.  A CTA owns a tile of 32 environments; lane l of EVERY warp stands for environment l
// of the tile and the CTA's R warps are ROLES: in each phase of the transition the independent
// work items of an environment (its cores, its job slots, its agents, its observation rows) are
// dealt out to the warps, so a warp runs one short, uniform instruction stream for 32
// environments (no role divergence inside a warp, an idle role costs no issue slots) and an
// environment's serial dependency chain is cut by the number of roles.  The phases meet at CTA
// barriers and exchange through the shared-memory tile:
//   P0  (while the TMA copies of the state/action tiles are in flight) Philox draws, zeroed
//       result record, observation background
//   P1  per core: who acts on it and which pending offer it selects (table index, or the
//       in-kernel auction: arg-max of offeredReward/necessaryTime by exact cross-multiplication,
//       random tie-break)
//   P2  one warp: executeAnOffer in reference order (agents asc, cores asc, auctioneer last)
//   P3  per core: job progress / completion, liability-chain walk, termination rewards
//   P4  per agent: spawn refill; per slot: offer creation
//   P5  scalar outputs; the new state and the result record leave with two bulk stores
//   P6  per core / per slot: observation rows of the new state, staged in the SAME shared memory
//       (the observation tile overlays the state/action/result tiles once the bulk stores have
//       read them: 12 KB per tile instead of 24, which lets every tile of a 65,536-env launch be
//       resident at once) and written with a third bulk store.
// R = 1 degenerates to one lane per environment with no cross-warp traffic at all.
#pragma once
#include "msched_common.cuh"
#include "observe_kernel.cuh"
#include "hardcoded_kernel.cuh"
#include "step_kernel.cuh"

namespace msched {

template <int N, int C, int L>
struct FusedDims {
    static constexpr int NL = N * L;
    static constexpr int NCH = (C + 3) / 4;
    static constexpr int W = (2 + 3 * C + NCH + 4 * NL) | 1;
    static constexpr int SCH = 2 + 3 * C;
    static constexpr int SSLOT = SCH + NCH;
    // per-env scratch words: sel[C] | key[C] | term[C] | tie draws 4*NCH | spawn draws 4 | #spawned
    // | occupancy mask of the slots before the spawn (2 words)
    static constexpr int X_SEL = 0, X_KEY = C, X_TERM = 2 * C, X_TIE = 3 * C, X_SPAWN = 3 * C + 4 * NCH;
    static constexpr int X_NSPAWN = X_SPAWN + 4, X_OCC = X_NSPAWN + 1;
    static constexpr int XW = (X_OCC + 2) | 1;
    // observation tile of 32 envs (bytes): small domains keep every tile of a big launch resident,
    // which needs 16 CTAs of <= 64 threads (8 of 128) per SM, i.e. at most 64 registers per thread
    static constexpr int OBS_TILE = 64 * ((N * C + C) * (2 * NL + 4) + NL * (2 * C + 2));
    static constexpr bool SMALL = OBS_TILE <= 14 * 1024;
};

inline size_t fused_smem_bytes(int stateWords, int actionHalfs, int resultWords, int obsHalfs, int C, int compactWords = 0)
{
    const int xw = (3 * C + 4 * ((C + 3) / 4) + 4 + 1 + 2) | 1;
    // the compact result tile (msched_step_host_compact) sits behind the work tiles
    const size_t work = (size_t)stateWords + actionHalfs / 2 + resultWords + xw + compactWords, obs = (size_t)obsHalfs / 2;
    return (size_t)32 * 4 * (work > obs ? work : obs);  // the observation tile overlays the work tiles
}

// one value of the compact result record: float planes as IEEE half (exact for the small integers and the 0.5 the
// reward functions produce; the host checks the domain), integer planes saturated to int16 (flag bit 4 if it bites)
__device__ __forceinline__ uint16_t compact_f(float v) { return __half_as_ushort(__float2half_rn(v)); }
__device__ __forceinline__ uint16_t compact_i(int v, uint32_t &flags)
{
    if (v > 32767 || v < -32768) { flags |= MSCHED_FLAG_COMPACT_RANGE; v = v > 0 ? 32767 : -32768; }
    return (uint16_t)(v & 0xffff);
}

// integer accumulation into the tile: several role warps may touch one word (exact, order-free
// atomics); a one-warp tile owns its words
template <int R>
__device__ __forceinline__ void acc_add(int *p, int v)
{
    if (R == 1) *p += v;
    else atomicAdd(p, v);
}
template <int R>
__device__ __forceinline__ void acc_or(uint32_t *p, uint32_t v)
{
    if (R == 1) *p |= v;
    else atomicOr(p, v);
}
template <int R>
__device__ __forceinline__ void acc_and(uint32_t *p, uint32_t v)
{
    if (R == 1) *p &= v;
    else atomicAnd(p, v);
}

// SPEC >= 0 fixes the configuration switches at compile time (the instantiations the BASELINE configs run):
// bits 0-1 reward variant, 2-3 auction mode, 4-5 spawn mode, bit 6 newJobsPerRoundPerAgent == 1
__host__ __device__ constexpr int fused_spec(int mode, int auction, int spawn, int newJobsIsOne)
{
    return mode | (auction << 2) | (spawn << 4) | (newJobsIsOne << 6);
}

template <int N, int C, int L, int R, int SPEC = -1, bool MULTI = false, bool HC = false>
// Resident CTAs per SM the registers are held to (small domains): 16 of 64 threads for the one-step kernel -- its
// launches overlap with their neighbours' and every resident tile counts (64 registers) -- and 14 for the multi-step
// kernel, whose one wave of 13.8 tiles per SM at 65,536 environments stays for the whole launch: 72 registers measured
// 15.2 us per step against 16.05 us (and the one-step kernel 16.5 against 15.9 us with them)
__global__ void __launch_bounds__(32 * R, FusedDims<N, C, L>::SMALL ? (R <= 2 ? (MULTI ? 14 : 16) : (MULTI ? 7 : 8)) : 1)
    fused_step_kernel(const __grid_constant__ DevParams p)
{
    using D = FusedDims<N, C, L>;
    constexpr int NL = D::NL, NCH = D::NCH, W = D::W, SCH = D::SCH, SSLOT = D::SSLOT, XW = D::XW;
    constexpr int RAw = NL + 2, ROw = C + 1;
    static_assert(C <= 8 && NL <= 64, "packed per-core counters / unrolled sweeps");
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int w = R == 1 ? 0 : (int)(threadIdx.x >> 5);  // compile-time for the one-warp tile: every role test folds
    const int env0 = blockIdx.x * 32, env = env0 + lane;
    const int AW = p.AH >> 1, RW = p.RW;
    const bool withObs = p.obs != nullptr;
    const int OW = withObs ? (p.OH >> 1) : 0;
    uint32_t *sState = sm, *sAct = sState + 32 * W, *sRes = sAct + 32 * AW, *sScr = sRes + 32 * RW;
    uint32_t *sObs = sm;  // overlays the work tiles in P6

    // per-CTA phase stamps (tools/timeline.py): compiled in only with -DMSCHED_TIMELINE, the shipped kernel
    // carries no diagnostic branches
#ifdef MSCHED_TIMELINE
    unsigned long long *tl = p.timeline ? p.timeline + (size_t)blockIdx.x * 8 : nullptr;
#define MSCHED_TL(stmt) do { if (tl && threadIdx.x == 0) { stmt; } } while (0)
#else
#define MSCHED_TL(stmt) do { } while (0)
#endif
    if (threadIdx.x == 0) {
        MSCHED_TL((tl[0] = smid(), tl[1] = globaltimer(), tl[2] = clock64()));
        mbar_init(&bar, 1);
    }

    const int auctionMode = SPEC >= 0 ? ((SPEC >> 2) & 3) : p.auctionMode;
    const int spawnMode = SPEC >= 0 ? ((SPEC >> 4) & 3) : p.spawnMode;
    const int newJobs = (SPEC >= 0 && (SPEC & 64)) ? 1 : p.newJobs;
    const bool randomTies = auctionMode == MSCHED_AUCTION_RANDOM_MAX;
    const bool external = auctionMode == MSCHED_AUCTION_EXTERNAL;
    const int mode = SPEC >= 0 ? (SPEC & 3) : p.mode;
    const bool agg = mode == MSCHED_REWARD_AGGREGATED_FIXED;
    const bool freeM = mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL || mode == MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL;
    const int round0 = cur_round(p);
    const bool live = env < p.B;
    // msched_step_multi: nSteps consecutive steps in this launch.  With observations after every step the tile is
    // re-read each step (the observation tile overlays it); without, the state tile stays in shared memory and only
    // the action tile of the next step is fetched
    const int nSteps = MULTI ? (p.nSteps > 1 ? p.nSteps : 1) : 1;  // (the one-step instantiation folds the loop away)
    constexpr bool hc = MULTI && HC;  // the hard-coded agents act inside the loop (msched_rollout_hardcoded): own instantiation
    constexpr int HCU = (N * C + NL + R - 1) / R;  // their units per thread
    unsigned ticket = 0u;
    bool stateResident = false, actResident = false;
#pragma unroll 1
    for (int tStep = 0; tStep < nSteps; ++tStep) {
    const int round = round0 + tStep;
    const bool lastStep = tStep == nSteps - 1;
    const bool obsStore = withObs && (lastStep || p.obsEvery != 0);
    const bool obsThis = obsStore || (withObs && hc);  // (the agents read the observation tile of every step)
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 32u * (uint32_t)((stateResident ? 0 : W) + (actResident ? 0 : AW)) * 4u);
        if (!stateResident) bulk_g2s(sState, p.state + (size_t)env0 * W, 32u * W * 4u, &bar);
        if (!actResident) bulk_g2s(sAct, p.action + (size_t)tStep * p.actStep + (size_t)env0 * p.AH, 32u * (uint32_t)AW * 4u, &bar);
    }

    uint32_t *st = sState + (size_t)lane * W;
    const int16_t *act = reinterpret_cast<const int16_t *>(sAct + (size_t)lane * AW);
    uint32_t *res = sRes + (size_t)lane * RW;
    uint32_t *scr = sScr + (size_t)lane * XW;
    float *resf = reinterpret_cast<float *>(res);
    int *resi = reinterpret_cast<int *>(res);
    uint32_t *core = st + 2, *chlS = st + SCH, *slot = st + SSLOT;

    // ---- P0: work that does not depend on the tile, overlapped with the copy ----
    if (w == 0 % R && randomTies) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            uint32_t x[4];
            env_draw_at(p, round, env, kStreamTie, (uint32_t)c, 0u, x);
            scr[D::X_TIE + 4 * c] = x[0]; scr[D::X_TIE + 4 * c + 1] = x[1];
            scr[D::X_TIE + 4 * c + 2] = x[2]; scr[D::X_TIE + 4 * c + 3] = x[3];
        }
    }
    if (w == 1 % R && spawnMode == MSCHED_SPAWN_PHILOX) {
        uint32_t x[4];
        env_draw_at(p, round, env, kStreamSpawn, 0u, 0u, x);
        scr[D::X_SPAWN] = x[0]; scr[D::X_SPAWN + 1] = x[1]; scr[D::X_SPAWN + 2] = x[2]; scr[D::X_SPAWN + 3] = x[3];
    }
    if (w == 2 % R) {
#pragma unroll 4
        for (int k = 0; k < RW; ++k) res[k] = 0u;
    }
    if (w == 3 % R) {
#pragma unroll
        for (int j = 0; j < C; ++j) scr[D::X_TERM + j] = 0u;
        scr[D::X_NSPAWN] = 0u;
    }
    if (round < 0) __trap();  // (never: makes every thread HOLD the round before the barrier, see take_round_ticket)
    __syncthreads();  // barrier initialisation and the P0 products visible to every warp
    if (tStep == 0 && threadIdx.x == 0) ticket = take_round_ticket(p);
    MSCHED_TL(tl[3] = clock64());
    mbar_wait(&bar, (uint32_t)tStep & 1u);
    MSCHED_TL(tl[4] = clock64());

    // ---- P1: per core, who acts on it and which pending offer is selected.  Offers addressed to
    // (owner, core) are ranked in creation order (agent asc, slot asc); agents (and an external
    // auctioneer) select by table index, the in-kernel auctioneer takes the arg-max of
    // offeredReward/necessaryTime compared exactly by cross-multiplication, -1/-2 operands rating
    // -1 (HardcodedModules.calculateRewardRatio) ----
    constexpr int CPW = (C + 1 + R - 1) / R;  // items per warp: the C cores + the "no core" item of P3
    uint4 pre[CPW];
    int preLen[CPW];
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        const int j = w + i * R;
        pre[i] = make_uint4(0u, 0u, 0u, 0u);
        preLen[i] = -1;
        if (j < C && live) {
            const uint32_t c0 = core[3 * j];
            const int o = core_owner(c0);
            // liability chain of a core whose job completes this step: issue the cold load now
            if (job_kind(c0) >= 0 && job_rem(c0) == 1) {
                const uint2 *cb = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
                if ((p.chainCap & 1) == 0) {
                    pre[i] = *reinterpret_cast<const uint4 *>(cb);
                    preLen[i] = (int)((chlS[j >> 2] >> ((j & 3) * 8)) & 0xffu);
                } else {
                    prefetch_l1(cb);
                }
            }
            // range check of the acceptor actions of this core (assert in src/world.py:389,404)
            bool bad = false;
#pragma unroll
            for (int a = 0; a < N; ++a) bad |= (unsigned)(int)act[p.aAcc + a * C + j] > (unsigned)NL;
            if (external) bad |= (unsigned)(int)act[p.aAuc + j] > (unsigned)NL;
            if (bad) acc_or<R>(&st[1], MSCHED_FLAG_ACTION_RANGE);

            int k = -1;
            if (o > 0) k = act[p.aAcc + (o - 1) * C + j];
            else if (external) k = act[p.aAuc + j];
            const bool auct = (o == 0) && !external;
            const int kk = (k >= 0 && k < NL) ? k : -1;
            const uint32_t match = (uint32_t)(j + 1) | ((uint32_t)o << 8);
            uint32_t sw0[NL], sw3[NL];
#pragma unroll
            for (int s = 0; s < NL; ++s) { sw0[s] = slot[4 * s]; sw3[s] = slot[4 * s + 3]; }
            int se = -1, sr = 0, cnt = 0, bn = -1, bd = 1, nc = 0;
#pragma unroll
            for (int s = 0; s < NL; ++s) {
                const uint32_t w3 = sw3[s];
                const bool m = (w3 & 0xffffu) == match;
                int pn = off_price(w3), pd = job_rem(sw0[s]);
                if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                const int lhs = pn * bd, rhs = bn * pd;
                const bool gt = m && auct && lhs > rhs;
                const bool eq = m && auct && lhs == rhs;
                const bool hit = (m && !auct && cnt == kk) || gt;
                se = hit ? s : se;
                sr = hit ? cnt : sr;
                bn = gt ? pn : bn;
                bd = gt ? pd : bd;
                nc = gt ? 1 : (eq ? nc + 1 : nc);
                cnt += m ? 1 : 0;
            }
            // uniformly random arg-max (random.sample in the reference): core j uses word j%4 of
            // Philox call j/4 of the tie stream
            if (auct && se >= 0 && nc > 1 && randomTies) {
                const int pick = (int)__umulhi(scr[D::X_TIE + j], (uint32_t)nc);
                if (pick > 0) {
                    int t = 0, rk = 0;
#pragma unroll
                    for (int s = 0; s < NL; ++s) {
                        const uint32_t w3 = sw3[s];
                        const bool m = (w3 & 0xffffu) == match;
                        int pn = off_price(w3), pd = job_rem(sw0[s]);
                        if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                        const bool tie = m && (pn * bd == bn * pd);
                        const bool hit = tie && t == pick;
                        se = hit ? s : se;
                        sr = hit ? rk : sr;
                        t += tie ? 1 : 0;
                        rk += m ? 1 : 0;
                    }
                }
            }
            scr[D::X_SEL + j] = (uint32_t)se;
            scr[D::X_KEY + j] = se >= 0 ? (uint32_t)((((o == 0) ? (N + 1) : o) << 8) | j) : 0x7fffffffu;
            const int kUsed = external ? (int)act[p.aAuc + j] : ((auct && se >= 0) ? sr : NL);
            reinterpret_cast<uint16_t *>(res + p.rAucIdx)[j] = (uint16_t)kUsed;
        }
    }
    __syncthreads();

    // ---- P2 (one warp): executeAnOffer in reference order: agents ascending, cores ascending,
    // auctioneer last (at most one acceptance per core, invariant I2) ----
    double qualSum = 0.0;
    int qualCnt = 0, nAcc = 0;
    if (w == 0 && live) {
        int key[C], sel[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { key[j] = (int)scr[D::X_KEY + j]; sel[j] = (int)scr[D::X_SEL + j]; }
        uint32_t flags = 0u;
        int lastKey = -1;
#pragma unroll
        for (int e = 0; e < C; ++e) {
            int best = 0x7fffffff;
#pragma unroll
            for (int jj = 0; jj < C; ++jj) best = (key[jj] > lastKey && key[jj] < best) ? key[jj] : best;
            if (best == 0x7fffffff) break;
            lastKey = best;
            const int j = best & 0xff;
            const int who = (best >> 8) == N + 1 ? 0 : (best >> 8);
            int se = sel[0];
#pragma unroll
            for (int jj = 1; jj < C; ++jj) se = (j == jj) ? sel[jj] : se;
            const int selA = se / L;
            const uint32_t c0 = core[3 * j], c1 = core[3 * j + 1], c2 = core[3 * j + 2];
            const uint32_t s0 = slot[4 * se], s1 = slot[4 * se + 1], s2 = slot[4 * se + 2], s3 = slot[4 * se + 3];
            const int kind = job_kind(s0), time = job_rem(s0), price = off_price(s3);
            const int offerer = selA + 1;
            const int prio1 = p.prio[kind];
            stat_accept(p, env, kind, price);
            slot[4 * se] = kEmptyJobW0; slot[4 * se + 1] = kEmptyId; slot[4 * se + 2] = kEmptyId; slot[4 * se + 3] = 0u;
            core[3 * j] = pack_core(offerer, kind, time);
            core[3 * j + 1] = s1;
            core[3 * j + 2] = s2;
            if (who > 0) {
                // old job back into the recipient's first empty slot
                const int base = (who - 1) * L;
                int q = -1;
#pragma unroll
                for (int t = L - 1; t >= 0; --t)
                    if (job_kind(slot[4 * (base + t)]) < 0) q = t;
                if (q >= 0) {
                    slot[4 * (base + q)] = c0 & 0xffffff00u;
                    slot[4 * (base + q) + 1] = c1;
                    slot[4 * (base + q) + 2] = c2;
                    slot[4 * (base + q) + 3] = 0u;
                } else {
                    flags |= MSCHED_FLAG_COLLECTION_FULL;
                }
                // acception quality, src/SchedulingEnvironment.py:174-192 (former = core before)
                double qv = __dmul_rn((double)price, c_rcp[time & 0xff]);
                const int fk = job_kind(c0);
                if (fk >= 0) qv = __dsub_rn(qv, __dmul_rn((double)p.prio[fk], c_rcp[job_rem(c0) & 0xff]));
                qualSum = __dadd_rn(qualSum, __dmul_rn(qv, 10.0));
                ++qualCnt;
            }
            {  // liability chain append (stored oldest first)
                const uint32_t cw = chlS[j >> 2];
                const int len = (int)((cw >> ((j & 3) * 8)) & 0xffu);
                if (len < p.chainCap) {
                    uint2 *ce = reinterpret_cast<uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap + len;
                    *ce = make_uint2((uint32_t)round, pack_chain(price, time, offerer));
                    chlS[j >> 2] = cw + (1u << ((j & 3) * 8));
                } else {
                    flags |= MSCHED_FLAG_CHAIN_OVERFLOW;
                }
            }
            if (agg) {
                resf[p.rOffer + selA] += (float)prio1;
            } else {
                resf[p.rOffer + se] = (float)prio1;
                if (freeM) {
                    const int df = prio1 - price;
                    float pr;
                    if (mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL)
                        pr = (df == 0) ? p.netZero : (float)df;
                    else
                        pr = (df >= 0) ? (float)prio1 : (float)df;
                    resf[p.rPrice + se] = pr;
                }
            }
            ++nAcc;
        }
        if (flags) acc_or<R>(&st[1], flags);
    }
    __syncthreads();

    // ---- P3: per core, job progress / completion + termination rewards (integer rewards shared
    // between cores are accumulated with shared-memory atomics: exact, order-free) ----
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        const int j = w + i * R;
        if (j < C && live) {
            const uint32_t c0 = core[3 * j];
            const int kind = job_kind(c0);
            if (kind >= 0) {
                const int rem = job_rem(c0) - 1;
                if (rem != 0) {
                    core[3 * j] = (c0 & 0x0000ffffu) | ((uint32_t)rem << 16);
                } else {
                    const int R_ = p.mult * p.prio[kind];
                    const int o = core_owner(c0) - 1;
                    stat_terminate(p, env, kind, round, core[3 * j + 2]);
                    if (agg) {
                        acc_add<R>(&resi[p.rAcc + o], R_);
                        acc_add<R>(&resi[p.rAgent + o], R_);
                    } else {
                        acc_add<R>(&resi[p.rAcc + o * C + j], R_);  // the only "=" into this word
                        if (!freeM) acc_add<R>(&resi[p.rAgent + o], R_);
                    }
                    const uint32_t cw = chlS[j >> 2];
                    const int len = (int)((cw >> ((j & 3) * 8)) & 0xffu);
                    const bool preOk = preLen[i] == len;  // an append in P2 makes the preload stale
                    const uint2 *ce = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
                    int recip = 0;  // the oldest entry was accepted by the auctioneer
                    for (int e = 0; e < len; ++e) {
                        uint2 en;
                        if (e < 2 && preOk) {
                            const uint4 q4 = pre[i];
                            en = e == 0 ? make_uint2(q4.x, q4.y) : make_uint2(q4.z, q4.w);
                        } else {
                            en = ce[e];
                        }
                        const int price = (int)(int16_t)(en.y & 0xffffu);
                        const int time = (int)((en.y >> 16) & 0xffu);
                        const int offerer = (int)(en.y >> 24);
                        const int traded = traded_reward(price, time, (round + 1) - (int)en.x);
                        acc_add<R>(&resi[p.rAgent + offerer - 1], -traded);
                        if (agg) {
                            acc_add<R>(&resi[p.rAcc + offerer - 1], -traded);
                            if (recip > 0) acc_add<R>(&resi[p.rAgent + recip - 1], traded);
                        } else {
                            acc_add<R>(&resi[p.rAcc + (offerer - 1) * C + j], -traded);
                            if (recip > 0) {
                                acc_add<R>(&resi[p.rAcc + (recip - 1) * C + j], traded);
                                acc_add<R>(&resi[p.rAgent + recip - 1], traded);
                            }
                        }
                        if (recip == 0) resi[p.rAuc + j] = traded;
                        recip = offerer;
                    }
                    acc_and<R>(&chlS[j >> 2], ~(0xffu << ((j & 3) * 8)));
                    core[3 * j] = kEmptyJobW0;
                    core[3 * j + 1] = kEmptyId;
                    core[3 * j + 2] = kEmptyId;
                    scr[D::X_TERM + j] = 1u;
                }
            }
        }
        // item j == C: which slots hold a job BEFORE the spawn (offers are created before it,
        // src/world.py:326-331, so a job spawned this step gets none)
        if (j == C && live) {
            unsigned long long occ = 0ull;
#pragma unroll
            for (int s = 0; s < NL; ++s) occ |= (unsigned long long)(job_kind(slot[4 * s]) >= 0 ? 1u : 0u) << s;
            scr[D::X_OCC] = (uint32_t)occ;
            scr[D::X_OCC + 1] = (uint32_t)(occ >> 32);
        }
    }
    __syncthreads();

    // ---- P4: spawn refill, one agent per item (src/world.py:369-376, src/Agent.py:50-70), and
    // offer creation, one slot per item (src/world.py:406-478).  They touch disjoint slots: an offer
    // needs a slot that was occupied before the spawn, the spawn fills empty ones ----
    if (live) {
        const unsigned long long occ = (unsigned long long)scr[D::X_OCC] | ((unsigned long long)scr[D::X_OCC + 1] << 32);
#pragma unroll
        for (int it0 = 0; it0 < NL; it0 += R) {
            const int s = it0 + (w + R - (N % R)) % R;  // slot items start on the warp after the last agent item
            if (s < NL && ((occ >> s) & 1ull)) {
                const int a = act[p.aOffc + s];
                const int kind = job_kind(slot[4 * s]);
                const bool waitOld = (slot[4 * s + 3] & 0xffu) != 0u;
                uint32_t w3 = 0u;
                if (a >= 0 && a < C && !waitOld) {
                    const int price = freeM ? (int)act[p.aOffp + s] : p.fix[kind];
                    w3 = pack_offer(a + 1, core_owner(core[3 * a]), price);
                }
                slot[4 * s + 3] = w3;
            }
        }
        for (int a = w; a < N; a += R) {
            // jobIDs are handed out in agent order: count the spawns of the lower agents
            uint32_t before = 0u;
            bool mine = false;
#pragma unroll
            for (int a2 = 0; a2 < N; ++a2) {
                if (a2 <= a) {
                    int owned = 0, nfree = 0;
#pragma unroll
                    for (int j = 0; j < C; ++j) owned += (core_owner(core[3 * j]) == a2 + 1) ? 1 : 0;
                    // free slots BEFORE any spawn of this step (another item may be filling some now)
                    nfree = L - __popcll((occ >> (a2 * L)) & ((1ull << L) - 1ull));
                    const bool sp = owned + newJobs <= nfree;
                    if (a2 < a) before += sp ? (uint32_t)newJobs : 0u;
                    else mine = sp;
                }
            }
            if (!mine) continue;
            uint32_t jobctr = st[0] + before;
            acc_add<R>(reinterpret_cast<int *>(&scr[D::X_NSPAWN]), newJobs);
            uint32_t rnd[4] = {scr[D::X_SPAWN], scr[D::X_SPAWN + 1], scr[D::X_SPAWN + 2], scr[D::X_SPAWN + 3]};
            int rndCall = 0;
            for (int k = 0; k < newJobs; ++k) {
                int kind = -1;
                if (spawnMode == MSCHED_SPAWN_KINDS) {
                    kind = act[p.aSpawn + a * newJobs + k];
                } else {
                    if (spawnMode == MSCHED_SPAWN_U64) {
                        const double u = p.spawnU[((size_t)env * N + a) * newJobs + k];
                        for (int q = 0; q < p.J; ++q)
                            if (u < p.cum[q]) { kind = q; break; }
                    } else {
                        const int dnum = a * newJobs + k;  // draw d uses word d%4 of Philox call d/4
                        if ((dnum >> 2) != rndCall) {
                            rndCall = dnum >> 2;
                            env_draw_at(p, round, env, kStreamSpawn, (uint32_t)rndCall, 0u, rnd);
                        }
                        const uint32_t xr = (dnum & 3) == 0 ? rnd[0] : (dnum & 3) == 1 ? rnd[1] : (dnum & 3) == 2 ? rnd[2] : rnd[3];
                        // first kind with u = xr * 2^-32 < cum[kind], decided exactly in integers
                        int cnt = 0;
                        for (int q = 0; q < p.J; ++q) cnt += ((unsigned long long)xr >= p.cumThr[q]) ? 1 : 0;
                        kind = cnt < p.J ? cnt : -1;
                    }
                }
                if (kind < 0 || kind >= p.J) { acc_or<R>(&st[1], MSCHED_FLAG_SPAWN_RANGE); kind = p.J - 1; }
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
    }
    __syncthreads();

    // ---- P5: job counter, scalar outputs, observations of the new state
    // (src/Agent.py:148-300, src/Auctioneer.py:20-77) ----
    if (live) {
        if (w == 0) {
            st[0] += scr[D::X_NSPAWN];  // the per-world job counter: every spawning agent took newJobs ids
            int nTerm = 0;
#pragma unroll
            for (int j = 0; j < C; ++j) nTerm += (int)scr[D::X_TERM + j];
            const uint32_t flags = st[1];
            const unsigned long long qb = (unsigned long long)__double_as_longlong(qualSum);
            res[p.rQual] = (uint32_t)qb;
            res[p.rQual + 1] = (uint32_t)(qb >> 32);
            res[p.rCounts] = (uint32_t)qualCnt | ((uint32_t)nAcc << 8) | ((uint32_t)nTerm << 16) | ((uint32_t)(nSteps > 1 ? ((((round + 1) % p.episodeLength) == 0) ? 1 : 0) : cur_done(p, round)) << 24);
            res[p.rFlags] = flags;
        }
    }
    // compact result record (host-buffer step): int16 / half planes, packed behind the work tiles
    if (p.cres && w == 0) {
        uint32_t *cr = sScr + 32 * XW + (size_t)lane * p.CW;
        uint16_t *ch = reinterpret_cast<uint16_t *>(cr);
        uint32_t cflags = 0u;
        const int nOff = N * p.RL, nAccW = N * p.RC;
        for (int k = 0; k < nOff; ++k) ch[p.cOffer + k] = live ? compact_f(resf[p.rOffer + k]) : (uint16_t)0;
        if (p.cPrice >= 0)
            for (int k = 0; k < nOff; ++k) ch[p.cPrice + k] = live ? compact_f(resf[p.rPrice + k]) : (uint16_t)0;
        for (int k = 0; k < nAccW; ++k) ch[p.cAcc + k] = live ? compact_i(resi[p.rAcc + k], cflags) : (uint16_t)0;
#pragma unroll
        for (int k = 0; k < C; ++k) ch[p.cAuc + k] = live ? compact_i(resi[p.rAuc + k], cflags) : (uint16_t)0;
#pragma unroll
        for (int k = 0; k < N; ++k) ch[p.cAgent + k] = live ? compact_i(resi[p.rAgent + k], cflags) : (uint16_t)0;
        for (int k = p.cAgent + N; k < 2 * p.cTail; ++k) ch[k] = 0;
        cr[p.cTail] = live ? __float_as_uint((float)qualSum) : 0u;
        cr[p.cTail + 1] = live ? ((res[p.rCounts] & 0x01ffffffu) | ((res[p.rFlags] | cflags) << 25)) : 0u;  // flags: bits 25..31
        for (int k = p.cTail + 2; k < p.CW; ++k) cr[k] = 0u;
    }

    // ---- the new state and the result record leave; the observation rows are then staged in the
    // same shared memory ----
    uint32_t cw0[C], sw0[NL], sw3[NL];
    if (obsThis) {
#pragma unroll
        for (int j = 0; j < C; ++j) cw0[j] = core[3 * j];
#pragma unroll
        for (int s = 0; s < NL; ++s) { sw0[s] = slot[4 * s]; sw3[s] = slot[4 * s + 3]; }
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        MSCHED_TL(tl[5] = clock64());
        // the state leaves when the observation tile is about to overlay it, and after the last step
        if (obsThis || lastStep) bulk_s2g(p.state + (size_t)env0 * W, sState, 32u * W * 4u);
        if (p.cres) bulk_s2g(p.cres + (size_t)env0 * p.CW, sScr + 32 * XW, 32u * (uint32_t)p.CW * 4u);
        else bulk_s2g(p.result + (size_t)tStep * p.resStep + (size_t)env0 * RW, sRes, 32u * (uint32_t)RW * 4u);
        bulk_commit();
        // the next step re-reads what this one wrote (state) only after the store has COMPLETED
        if (obsThis && !lastStep) bulk_wait_all();
        else bulk_wait_read();
    }
    stateResident = !obsThis;
    if (obsThis) {
    __syncthreads();  // the bulk stores have read the work tiles: the memory is free

    // ---- P6: observations of the new state (src/Agent.py:148-300, src/Auctioneer.py:20-77).  A core
    // item writes all N+1 rows of its core (the owner's, or the auctioneer's, carries the core's job
    // and the offers addressed to it; the others are the constant [0,-1,-1,-2,...]); a slot item
    // writes the slot's offer-net row ----
    {
        uint32_t *ob = sObs + (size_t)lane * OW;
        uint32_t cp[C];
#pragma unroll
        for (int j = 0; j < C; ++j) cp[j] = job_pair(p, cw0[j]);
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int j = w + i * R;
            if (j < C) {
#pragma unroll
                for (int a = 0; a <= N; ++a) {
                    uint32_t *row = ob + (a * C + j) * RAw;
                    row[0] = 0u;
                    row[1] = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < NL; ++k) row[2 + k] = 0xfffefffeu;
                }
                if (live) {
                    const int o = core_owner(cw0[j]);
                    uint32_t *row = ob + (o > 0 ? (o - 1) * C + j : N * C + j) * RAw;
                    row[0] = 0x00010000u;  // [pad, own = 1]
                    row[1] = cp[j];
                    // offers to core j carry recipient == its owner: creation order = slot order
                    int n = 0;
#pragma unroll
                    for (int s = 0; s < NL; ++s) {
                        if ((int)(sw3[s] & 0xffu) == j + 1) {
                            row[2 + n] = pair16(off_price(sw3[s]), job_rem(sw0[s]));
                            ++n;
                        }
                    }
                }
            }
        }
        uint32_t *oo = ob + (N * C + C) * RAw;
#pragma unroll
        for (int s0 = 0; s0 < NL; s0 += R) {
            const int s = s0 + w;
            if (s < NL) {
#pragma unroll
                for (int j = 0; j < C; ++j) oo[s * ROw + j] = cp[j];
                oo[s * ROw + C] = job_pair(p, sw0[s]);
            }
        }
        if (w == R - 1)
            for (int k = (N * C + C) * RAw + NL * ROw; k < OW; ++k) ob[k] = 0u;
    }
    fence_async_smem();
    __syncthreads();
    // DividedHardcodedAgent.getActions on the observation tile (src/Agent.py:622-641, src/HardcodedModules.py:16-45,
    // 81-109): the units of the lane's environment are dealt to the role warps
    int16_t hcAct[HCU];
    if constexpr (hc) {
        const int16_t *ob16 = reinterpret_cast<const int16_t *>(sObs + (size_t)lane * OW);
#pragma unroll
        for (int i = 0; i < HCU; ++i) {
            const int unit = w + i * R;
            hcAct[i] = 0;
            if (unit < N * C + NL) {
                const float u = p.hcRandomTies ? hc_unit_draw(p, round + 1, env, unit) : 0.f;
                int nc;
                hcAct[i] = (int16_t)hc_unit_action(p, ob16, p.hcOAcc, p.hcOOff, p.hcAccRow, p.hcOffRow, unit, u, nc);
            }
        }
    }
    if (threadIdx.x == 0 && obsStore) {
        bulk_s2g(reinterpret_cast<uint32_t *>(p.obs + (p.obsEvery ? (size_t)tStep * p.obsStep : 0)) + (size_t)env0 * OW, sObs,
                 32u * (uint32_t)OW * 4u);
        bulk_commit();
        bulk_wait_read();
    }
    if constexpr (hc) {
        __syncthreads();  // the rows have been read (and the observation store has read the tile)
        // the next step's action tile, or, after the last step, the action record in global memory
        int16_t *dstA = lastStep ? p.actionOut + (size_t)env * p.AH : reinterpret_cast<int16_t *>(sAct + (size_t)lane * AW);
        if (!lastStep || live) {
#pragma unroll
            for (int i = 0; i < HCU; ++i) {
                const int unit = w + i * R;
                if (unit < N * C + NL) dstA[unit < N * C ? p.aAcc + unit : p.aOffc + unit - N * C] = hcAct[i];
            }
        }
        actResident = !lastStep;
    }
    }  // obsThis
    if (!lastStep) __syncthreads();  // the stores have read the tiles: the next step may overwrite them
    }  // steps of the launch
    if (threadIdx.x == 0) {
        MSCHED_TL((tl[6] = clock64(), tl[7] = globaltimer()));
        redeem_round_ticket(p, ticket, nSteps);
    }
#undef MSCHED_TL
}

}  // namespace msched
