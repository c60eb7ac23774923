// warp_step_kernel.cuh -- one SchedulingEnv.step with ONE WARP per environment, for the large domains
// (BASELINE config 5: N32 C64 L8 = 64 cores, 256 job slots, up to 256 pending offers per environment).
//
// Same reference semantics and record formats as step_kernel.cuh:
//   src/world.py:295-334                 World.step1 (order of phases)
//   src/world.py:261-293, 378-404        executeAnOffer, agent / auctioneer acceptances
//   src/HardcodedModules.py:5-13, 48-78  the auction: arg-max of offeredReward/necessaryTime, random tie
//   src/world.py:336-367, 406-478        progress/completion, offer creation
//   src/world.py:369-376, src/Agent.py:50-70   spawn refill
//   src/Reward.py:6-212                  rewards
//   src/SchedulingEnvironment.py:174-192 acception quality
//
// Mapping.  A CTA of four warps owns four consecutive environments; their state records are one
// contiguous 16*W-byte chunk that a single TMA bulk copy brings into shared memory (and takes back).
// Warp w advances environment w with all 32 lanes; every phase is O(slots/32 + cores/32) warp-wide
// steps instead of the O(slots) serial sweep of a lane- or group-per-environment kernel:
//   * the pending offers are swept 32 slots at a time in creation order (agent asc, slot asc).  The lanes
//     whose offers address the same core find each other with MATCH.ANY (__match_any_sync); an offer's
//     rank inside its (recipient, core) table is the core's running count plus the number of lower lanes in
//     its match group -- that rank is what an agent's acceptor index selects;
//   * the auction of an idle core is an arg-max over an exactly ordered 64-bit key, floor(price * 2^32 /
//     time): job lengths are <= 255 and prices int16, so two different ratios differ by more than 2^-16 and
//     their keys differ, equal ratios give equal keys (no floating point, no cross-multiplication chain).
//     The maximum is a shared-memory atomicMax per offer; a second sweep ranks the tied maxima with the
//     same match/ballot scheme and the Philox draw of the core picks one (random.sample in the reference);
//   * executeAnOffer is serial in reference order (agents asc, cores asc, auctioneer last) but each
//     execution is a warp-wide step: the next (owner, core) key is a REDUX min over the lanes' cores, the
//     recipient's first empty slot is a ballot + find-first-set;
//   * progress/completion: one lane per core; offer creation: one lane per slot; spawn refill: one lane
//     per agent, the cores an agent owns counted by a shared-memory histogram, job IDs by a ballot prefix.
// Nothing but the 4.9 KB state record (config 5) and ~2.5 KB of scratch lives in shared memory: the
// action record is read straight from global memory (coalesced, each value used once) and the result
// record -- 9.8 KB per environment in config 5, almost all zeros -- is zero-filled in global memory with
// 16-byte stores while the state tile is in flight, then patched with the few non-zero rewards.  That is
// what lets 28 environments be resident per SM instead of 8.  The compact observation record (cores,
// slots, pending offers as int16, 2.5 KB instead of the 2.2 MB dense record) leaves from the same launch.
#pragma once
#include "msched_common.cuh"
#include "reward_math.cuh"

namespace msched {

// floor(n / d) for 0 <= n < 2^24, 1 <= d <= 255 is umulhi(n, ceil(2^32 / d)) exactly: the error of the
// reciprocal adds less than 2^-8 to n/d and frac(n/d) <= 1 - 1/255 < 1 - 2^-8
__constant__ uint32_t c_magic[256];

// order-preserving key of calculateRewardRatio(price, time) (src/HardcodedModules.py:5-13): the -1 / -2
// paddings rate -1; otherwise floor(price * 2^32 / time), biased to unsigned
__device__ __forceinline__ unsigned long long ratio_key(int pn, int pd)
{
    if (pn == -1 || pn == -2 || pd == -1 || pd == -2 || pd <= 0 || pd > 255) { pn = -1; pd = 1; }
    const uint32_t m = c_magic[pd];
    int q;
    if (pd == 1) q = pn;
    else if (pn >= 0) q = (int)__umulhi((uint32_t)pn, m);
    else q = -(int)__umulhi((uint32_t)(-pn + pd - 1), m);
    const uint32_t r = (uint32_t)(pn - q * pd);  // 0 <= r < pd
    uint32_t lo = 0u;
    if (pd != 1) {
        const uint32_t t1 = __umulhi(r << 16, m), r1 = (r << 16) - t1 * (uint32_t)pd;
        const uint32_t t2 = __umulhi(r1 << 16, m);
        lo = (t1 << 16) | t2;
    }
    return (((unsigned long long)(uint32_t)q << 32) | lo) ^ 0x8000000000000000ull;
}
constexpr unsigned long long kRatioMinusOne = (0xffffffff00000000ull) ^ 0x8000000000000000ull;  // ratio_key(-1, 1)

// ---- compact observation record (int16 per env, msched_get_compact_layout) ----
//   core  [C][4]  ownerID (0 = auctioneer), priority, remainingLength, jobKind   (-1 = idle)
//   slot  [NL][2] priority, remainingLength                                       (-1 = empty)
//   offer [NL][2] coreID (0 = no pending offer), offeredReward; the recipient is that core's owner, the
//                 necessaryTime the slot's remainingLength (invariant I4)
// Everything the dense rows of src/Agent.py:167-300 / src/Auctioneer.py:34-77 are built from.
__host__ __device__ inline int compact_halfs(int C, int NL)
{
    int h = 4 * C + 4 * NL;
    if (((h / 2) & 1) == 0) h += 2;  // odd word count like the other records
    return h;
}
// st: the env's state record (shared or global memory); out: its compact record (global); all 32 lanes
__device__ __forceinline__ void emit_compact(const DevParams &p, const uint32_t *st, int16_t *out, int lane)
{
    const int C = p.C, NL = p.NL;
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
    const uint32_t *core = st + 2, *slot = st + p.sSlot;
    for (int j = lane; j < C; j += 32) {
        const uint32_t w0 = core[3 * j];
        const int kind = job_kind(w0);
        o[2 * j] = (uint32_t)core_owner(w0) | ((uint32_t)(kind >= 0 ? p.prio[kind] : -1) << 16);
        o[2 * j + 1] = (uint32_t)(job_rem(w0) & 0xffff) | ((uint32_t)(kind & 0xffff) << 16);
    }
    uint32_t *os = o + 2 * C, *oo = os + NL;
    for (int s = lane; s < NL; s += 32) {
        const uint32_t w0 = slot[4 * s], w3 = slot[4 * s + 3];
        const int kind = job_kind(w0);
        os[s] = (uint32_t)((kind >= 0 ? p.prio[kind] : -1) & 0xffff) | ((uint32_t)(job_rem(w0) & 0xffff) << 16);
        oo[s] = (w3 & 0xffu) ? ((w3 & 0xffu) | (w3 & 0xffff0000u)) : 0u;
    }
    for (int k = 2 * C + 2 * NL + lane; k < (p.COH >> 1); k += 32) o[k] = 0u;
}

// stand-alone compact observations of the current state (msched_observe_compact): one warp per env
__global__ void __launch_bounds__(128) observe_compact_kernel(const __grid_constant__ DevParams p)
{
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= p.B) return;
    emit_compact(p, p.state + (size_t)env * p.W, p.cobs + (size_t)env * p.COH, threadIdx.x & 31);
}

// per-warp scratch (bytes): int16 per core: owner, kidx, sel, selRank, cnt, tie, pick | u64 per core: best key
// | u16 per slot: rank, tieRank | int per agent: cores owned
__host__ __device__ inline int warp_scratch_bytes(int N, int C, int NL)
{
    const int b = ((7 * 2 * C + 7) & ~7) + 8 * C + 2 * 2 * NL + 4 * N;
    return (b + 15) & ~15;
}
inline size_t warp_step_smem(int W, int N, int C, int NL)
{
    return (size_t)16 * W + (size_t)4 * warp_scratch_bytes(N, C, NL);
}

__global__ void __launch_bounds__(128, 7) warp_step_kernel(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env0 = blockIdx.x * 4, env = env0 + warp;
    const int N = p.N, C = p.C, L = p.L, NL = p.NL, W = p.W, RW = p.RW;
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned ltMask = (1u << lane) - 1u;

    uint32_t *sState = reinterpret_cast<uint32_t *>(smem);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 16u * (uint32_t)W);
        bulk_g2s(sState, p.state + (size_t)env0 * W, 16u * (uint32_t)W, &bar);
    }
    // ---- P0 (the tile is in flight): zero the four result records, 16 bytes per store ----
    {
        uint4 *rz = reinterpret_cast<uint4 *>(p.result + (size_t)env0 * RW);  // 4*RW words: 16-byte aligned
        for (int k = threadIdx.x; k < RW; k += 128) rz[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    unsigned char *scrB = smem + (size_t)16 * W + (size_t)warp * warp_scratch_bytes(N, C, NL);
    int16_t *sOwner = reinterpret_cast<int16_t *>(scrB), *sKidx = sOwner + C, *sSel = sKidx + C, *sSelRank = sSel + C;
    uint16_t *sCnt = reinterpret_cast<uint16_t *>(sSelRank + C), *sTie = sCnt + C, *sPick = sTie + C;
    unsigned long long *sBest = reinterpret_cast<unsigned long long *>(scrB + ((7 * 2 * C + 7) & ~7));
    uint16_t *sRank = reinterpret_cast<uint16_t *>(sBest + C), *sTieRank = sRank + NL;
    int *sOwned = reinterpret_cast<int *>(sTieRank + NL);

    const bool live = env < p.B;
    const int mode = p.mode;
    const bool agg = mode == MSCHED_REWARD_AGGREGATED_FIXED;
    const bool freeM = mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL || mode == MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL;
    const bool external = p.auctionMode == MSCHED_AUCTION_EXTERNAL;
    const bool randomTies = p.auctionMode == MSCHED_AUCTION_RANDOM_MAX;
    const int round = cur_round(p);
    const int16_t *act = p.action + (size_t)env * p.AH;
    uint32_t flags = 0u;

    // range check of every acceptor action (assert in src/world.py:389,404): packed int16 min / max
    if (live) {
        const uint32_t *aw = reinterpret_cast<const uint32_t *>(act + p.aAcc);  // aAcc == 0: word aligned
        const int nc = N * C;
        uint32_t mx = 0u, mn = 0u;
        for (int k = lane; k < (nc >> 1); k += 32) {
            const uint32_t v = aw[k];
            mx = __vmaxs2(mx, v);
            mn = __vmins2(mn, v);
        }
        int hi = max((int)(int16_t)(mx & 0xffffu), (int)(int16_t)(mx >> 16));
        int lo = min((int)(int16_t)(mn & 0xffffu), (int)(int16_t)(mn >> 16));
        if ((nc & 1) && lane == 0) { const int v = act[p.aAcc + nc - 1]; hi = max(hi, v); lo = min(lo, v); }
        if (hi > NL || lo < 0) flags |= MSCHED_FLAG_ACTION_RANGE;
    }
    __syncthreads();  // barrier initialisation visible; the zero fill ordered before the reward patches
    mbar_wait(&bar, 0);

    uint32_t *st = sState + (size_t)warp * W;
    uint32_t *core = st + 2, *slot = st + p.sSlot;
    unsigned char *chl = reinterpret_cast<unsigned char *>(st + p.sChlen);
    uint32_t *res = p.result + (size_t)env * RW;
    float *resf = reinterpret_cast<float *>(res);
    int *resi = reinterpret_cast<int *>(res);
    int nAcc = 0, nTerm = 0, qualCnt = 0;
    double qualSum = 0.0;

    if (live) {
        // ---- P1a: per core, who acts on it and with which table index ----
        for (int j = lane; j < C; j += 32) {
            const uint32_t c0 = core[3 * j];
            const int o = core_owner(c0);
            int k = -1;
            if (o > 0) k = act[p.aAcc + (o - 1) * C + j];
            else if (external) {
                k = act[p.aAuc + j];
                if (k < 0 || k > NL) flags |= MSCHED_FLAG_ACTION_RANGE;
            }
            sOwner[j] = (int16_t)o;
            sKidx[j] = (int16_t)((k >= 0 && k < NL) ? k : -1);
            sSel[j] = -1; sSelRank[j] = 0; sCnt[j] = 0; sTie[j] = 0; sPick[j] = 0;
            sBest[j] = kRatioMinusOne;
            // liability chain of a core whose job completes this step: issue the cold load now
            if (job_kind(c0) >= 0 && job_rem(c0) == 1)
                prefetch_l1(reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap);
        }
        for (int a = lane; a < N; a += 32) sOwned[a] = 0;
        __syncwarp();

        // ---- P1b: first sweep over the pending offers, 32 slots at a time in creation order: rank inside the
        // (recipient, core) table, selection by the acceptor index, the auction's maximum ----
        bool anyAuction = false;
        for (int s0 = 0; s0 < NL; s0 += 32) {
            const int s = s0 + lane;
            const uint32_t w3 = s < NL ? slot[4 * s + 3] : 0u;
            const int c = (int)(w3 & 0xffu);
            const int j = c - 1;
            const int o = c ? (int)sOwner[j] : -1;
            const bool valid = c != 0 && off_recip(w3) == o;  // addressed to the core's owner, else in nobody's table
            const unsigned peers = __match_any_sync(FULL, valid ? j : (0x100 | lane));
            const int base = valid ? (int)sCnt[j] : 0;
            __syncwarp();
            if (valid && (peers & ltMask) == 0u) sCnt[j] = (uint16_t)(base + __popc(peers));
            const int rank = base + __popc(peers & ltMask);
            if (valid) {
                sRank[s] = (uint16_t)rank;
                if (o > 0 || external) {
                    if (rank == (int)sKidx[j]) { sSel[j] = (int16_t)s; sSelRank[j] = (int16_t)rank; }
                } else {
                    atomicMax(&sBest[j], ratio_key(off_price(w3), job_rem(slot[4 * s])));
                    anyAuction = true;
                }
            }
            __syncwarp();
        }
        // ---- P1c: the auction's winner among the tied maxima (uniformly random: random.sample) ----
        if (__any_sync(FULL, anyAuction)) {
            for (int s0 = 0; s0 < NL; s0 += 32) {
                const int s = s0 + lane;
                const uint32_t w3 = s < NL ? slot[4 * s + 3] : 0u;
                const int c = (int)(w3 & 0xffu);
                const int j = c - 1;
                bool tie = false;
                if (c != 0 && off_recip(w3) == 0 && sOwner[j] == 0 && !external) {
                    const unsigned long long b = sBest[j];
                    tie = b > kRatioMinusOne && ratio_key(off_price(w3), job_rem(slot[4 * s])) == b;
                }
                const unsigned peers = __match_any_sync(FULL, tie ? j : (0x100 | lane));
                const int base = tie ? (int)sTie[j] : 0;
                __syncwarp();
                if (tie && (peers & ltMask) == 0u) sTie[j] = (uint16_t)(base + __popc(peers));
                if (s < NL) sTieRank[s] = tie ? (uint16_t)(base + __popc(peers & ltMask)) : (uint16_t)0xffffu;
                __syncwarp();
            }
            if (randomTies) {
                for (int j = lane; j < C; j += 32) {
                    const int nc = sTie[j];
                    if (nc > 1) {  // core j uses word j%4 of Philox call j/4 of the tie stream
                        uint32_t x[4];
                        env_draw(p, env, kStreamTie, (uint32_t)(j >> 2), 0u, x);
                        const uint32_t xw = (j & 3) == 0 ? x[0] : (j & 3) == 1 ? x[1] : (j & 3) == 2 ? x[2] : x[3];
                        sPick[j] = (uint16_t)__umulhi(xw, (uint32_t)nc);
                    }
                }
                __syncwarp();
            }
            for (int s = lane; s < NL; s += 32) {
                const uint16_t tr = sTieRank[s];
                if (tr != 0xffffu) {
                    const int j = (int)(slot[4 * s + 3] & 0xffu) - 1;
                    if (tr == sPick[j]) { sSel[j] = (int16_t)s; sSelRank[j] = (int16_t)sRank[s]; }
                }
            }
            __syncwarp();
        }
        // the auctioneer's index per core, as Auctioneer.getAuctioneerAction reports it
        for (int j = lane; j < C; j += 32) {
            int kUsed = NL;
            if (external) kUsed = act[p.aAuc + j];
            else if (sOwner[j] == 0 && sSel[j] >= 0) kUsed = sSelRank[j];
            reinterpret_cast<uint16_t *>(res + p.rAucIdx)[j] = (uint16_t)kUsed;
        }

        // ---- P2: executeAnOffer in reference order: agents ascending, cores ascending, auctioneer last (at
        // most one acceptance per core, invariant I2); every execution is a warp-wide step ----
        int lastKey = -1;
        for (;;) {
            int m = 0x7fffffff;
            for (int j = lane; j < C; j += 32) {
                if (sSel[j] >= 0) {
                    const int o = sOwner[j];
                    const int key = (((o == 0) ? (N + 1) : o) << 8) | j;
                    if (key > lastKey && key < m) m = key;
                }
            }
            const int best = (int)__reduce_min_sync(FULL, (unsigned)m);
            if (best == 0x7fffffff) break;
            lastKey = best;
            const int j = best & 0xff;
            const int who = (best >> 8) == N + 1 ? 0 : (best >> 8);
            const int se = sSel[j];
            const int selA = se / L;
            const uint32_t c0 = core[3 * j], c1 = core[3 * j + 1], c2 = core[3 * j + 2];
            const uint32_t s0w = slot[4 * se], s1w = slot[4 * se + 1], s2w = slot[4 * se + 2], s3w = slot[4 * se + 3];
            const int kind = job_kind(s0w), time = job_rem(s0w), price = off_price(s3w);
            const int offerer = selA + 1;
            const int prio1 = p.prio[kind];
            __syncwarp();
            if (lane == 0) {
                stat_accept(p, env, kind, price);
                slot[4 * se] = kEmptyJobW0; slot[4 * se + 1] = kEmptyId; slot[4 * se + 2] = kEmptyId; slot[4 * se + 3] = 0u;
                core[3 * j] = pack_core(offerer, kind, time);
                core[3 * j + 1] = s1w;
                core[3 * j + 2] = s2w;
            }
            __syncwarp();
            if (who > 0) {
                // old job back into the recipient's first empty slot: ballot + find-first-set
                const int base = (who - 1) * L;
                int q = -1;
                for (int t0 = 0; t0 < L && q < 0; t0 += 32) {
                    const int t = t0 + lane;
                    const unsigned bal = __ballot_sync(FULL, t < L && job_kind(slot[4 * (base + t)]) < 0);
                    if (bal) q = t0 + __ffs(bal) - 1;
                }
                if (lane == 0) {
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
            }
            if (lane == 0) {
                const int len = chl[j];  // liability chain append (stored oldest first)
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
                    resf[p.rOffer + se] = (float)prio1;
                    if (freeM) {
                        const int df = prio1 - price;
                        float pr;
                        if (mode == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL) pr = (df == 0) ? p.netZero : (float)df;
                        else pr = (df >= 0) ? (float)prio1 : (float)df;
                        resf[p.rPrice + se] = pr;
                    }
                }
            }
            ++nAcc;
            __syncwarp();
        }

        // ---- P3: job progress / completion + termination rewards, one lane per core.  A reward word shared
        // between cores (agent totals) is accumulated with atomics: exact, order-free ----
        for (int j = lane; j < C; j += 32) {
            const uint32_t c0 = core[3 * j];
            const int kind = job_kind(c0);
            if (kind < 0) continue;
            const int rem = job_rem(c0) - 1;
            if (rem != 0) {
                core[3 * j] = (c0 & 0x0000ffffu) | ((uint32_t)rem << 16);
                atomicAdd(&sOwned[core_owner(c0) - 1], 1);
                continue;
            }
            const int R_ = p.mult * p.prio[kind];
            const int o = core_owner(c0) - 1;
            stat_terminate(p, env, kind, round, core[3 * j + 2]);
            if (agg) {
                atomicAdd(&resi[p.rAcc + o], R_);
                atomicAdd(&resi[p.rAgent + o], R_);
            } else {
                resi[p.rAcc + o * C + j] = R_;  // column j belongs to this lane
                if (!freeM) atomicAdd(&resi[p.rAgent + o], R_);
            }
            const int len = chl[j];
            const uint2 *ce = reinterpret_cast<const uint2 *>(p.chain) + ((size_t)env * C + j) * p.chainCap;
            int recip = 0;  // the oldest entry was accepted by the auctioneer
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
        __syncwarp();

        // ---- P4: offer creation, one lane per slot (src/world.py:406-478); jobs spawned below get none ----
        for (int s = lane; s < NL; s += 32) {
            const uint32_t w0 = slot[4 * s];
            const int kind = job_kind(w0);
            const bool waitOld = (slot[4 * s + 3] & 0xffu) != 0u;
            uint32_t w3 = 0u;
            if (kind >= 0 && !waitOld) {
                const int a = act[p.aOffc + s];
                if (a >= 0 && a < C) {
                    const int price = p.freePrices ? (int)act[p.aOffp + s] : p.fix[kind];
                    w3 = pack_offer(a + 1, core_owner(core[3 * a]), price);
                }
            }
            slot[4 * s + 3] = w3;
        }
        __syncwarp();

        // ---- P5: spawn refill, one lane per agent (src/world.py:369-376, src/Agent.py:50-70); job IDs are
        // handed out in agent order: ballot prefix ----
        uint32_t jobctr = st[0];
        for (int a0 = 0; a0 < N; a0 += 32) {
            const int a = a0 + lane;
            bool spawn = false;
            if (a < N) {
                int nfree = 0;
                for (int q = 0; q < L; ++q) nfree += job_kind(slot[4 * (a * L + q)]) < 0 ? 1 : 0;
                spawn = sOwned[a] + p.newJobs <= nfree;
            }
            const unsigned bal = __ballot_sync(FULL, spawn);
            if (spawn) {
                uint32_t id = jobctr + (uint32_t)(__popc(bal & ltMask) * p.newJobs);
                uint32_t rnd[4] = {0u, 0u, 0u, 0u};
                int rndCall = -1;
                for (int k = 0; k < p.newJobs; ++k) {
                    int kind = -1;
                    if (p.spawnMode == MSCHED_SPAWN_KINDS) {
                        kind = act[p.aSpawn + a * p.newJobs + k];
                    } else if (p.spawnMode == MSCHED_SPAWN_U64) {
                        const double u = p.spawnU[((size_t)env * N + a) * p.newJobs + k];
                        for (int q = 0; q < p.J; ++q)
                            if (u < p.cum[q]) { kind = q; break; }
                    } else {
                        const int dnum = a * p.newJobs + k;  // draw d uses word d%4 of Philox call d/4
                        if ((dnum >> 2) != rndCall) {
                            rndCall = dnum >> 2;
                            env_draw(p, env, kStreamSpawn, (uint32_t)rndCall, 0u, rnd);
                        }
                        const uint32_t xr = (dnum & 3) == 0 ? rnd[0] : (dnum & 3) == 1 ? rnd[1] : (dnum & 3) == 2 ? rnd[2] : rnd[3];
                        int cnt = 0;  // first kind with u = xr * 2^-32 < cum[kind], decided exactly in integers
                        for (int q = 0; q < p.J; ++q) cnt += ((unsigned long long)xr >= p.cumThr[q]) ? 1 : 0;
                        kind = cnt < p.J ? cnt : -1;
                    }
                    if (kind < 0 || kind >= p.J) { flags |= MSCHED_FLAG_SPAWN_RANGE; kind = p.J - 1; }
                    int q = 0;
                    for (int t = L - 1; t >= 0; --t)
                        if (job_kind(slot[4 * (a * L + t)]) < 0) q = t;
                    const int s = a * L + q;  // an empty slot exists by the guard above
                    slot[4 * s] = pack_slot(kind, p.len[kind]);
                    slot[4 * s + 1] = id++;
                    slot[4 * s + 2] = (uint32_t)round;
                    slot[4 * s + 3] = 0u;
                }
            }
            jobctr += (uint32_t)(__popc(bal) * p.newJobs);
        }

        // ---- P6: scalar outputs ----
        flags = __reduce_or_sync(FULL, flags);
        nTerm = (int)__reduce_add_sync(FULL, (unsigned)nTerm);
        if (lane == 0) {
            flags |= st[1];
            st[0] = jobctr;
            st[1] = flags;
            const unsigned long long qb = (unsigned long long)__double_as_longlong(qualSum);
            res[p.rQual] = (uint32_t)qb;
            res[p.rQual + 1] = (uint32_t)(qb >> 32);
            res[p.rCounts] = (uint32_t)qualCnt | ((uint32_t)nAcc << 8) | ((uint32_t)nTerm << 16) | ((uint32_t)cur_done(p, round) << 24);
            res[p.rFlags] = flags;
        }
        __syncwarp();
        // ---- P7: compact observations of the new state ----
        if (p.cobs) emit_compact(p, st, p.cobs + (size_t)env * p.COH, lane);
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_s2g(p.state + (size_t)env0 * W, sState, 16u * (uint32_t)W);
        bulk_commit();
        bulk_wait_read();
        finish_round(p);
    }
}

}  // namespace msched
