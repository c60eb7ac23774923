// observe_kernel.cuh -- dense reference-layout observations and the reference-shaped state dump.
//
//   src/Agent.py:148-212    DividedAgent.gatherObservations / getAcceptorObservationTensorAndIDs
//   src/Agent.py:271-300    getOfferNetObservationTensor
//   src/Auctioneer.py:20-77 gatherDividedAuctioneerObservation
//
// Observation record (int16 per env, offsets from msched_get_layout).  Row CONTENTS are the
// reference's; rows are padded so that every (value, value) pair is one aligned 32-bit word:
//   acceptor rows   [N*C] x RA halfs, RA = 3+2NL+1: [pad, own?1:0, own?prio:-1, own?rem:-1,
//                   (price,time) of the offers addressed to (agent, core) in creation order,
//                   (-2,-2) padding]; the logical row starts at +1 (o_acceptor is odd)
//   auctioneer rows [C] x RA, same with owner 0
//   offer rows      [N*L] x RO halfs, RO = 2C+2: (prio, rem) of every core, then of the slot
// The semi-/fully-aggregated layouts (src/Agent.py:82-140, 399-461) are concatenations of these
// rows and are assembled as views on the host side.  Offer-ID tables (only the drop-in API and
// the parity tests need them) are a separate record written by ids_kernel.
#pragma once
#include "msched_common.cuh"

namespace msched {

__device__ __forceinline__ uint32_t pair16(int lo, int hi)
{
    return (uint32_t)(lo & 0xffff) | ((uint32_t)(hi & 0xffff) << 16);
}
__device__ __forceinline__ uint32_t job_pair(const DevParams &p, uint32_t w0)
{  // (priority, remainingLength) of a core/slot job word; empty -> (-1, -1)
    const int kind = job_kind(w0);
    return kind >= 0 ? pair16(p.prio[kind], job_rem(w0)) : 0xffffffffu;
}

// ---- fast path: compile-time domain, one warp per 32-env tile --------------------------------
// The state tile lands (TMA) in the front of the staging buffer, every lane lifts its record into
// registers, and the same buffer is then overwritten with the observation tile that one bulk
// store writes back: shared memory per env = the observation record only.
template <int N, int C, int L>
__global__ void __launch_bounds__(32) observe_kernel_t(const __grid_constant__ DevParams p)
{
    constexpr int NL = N * L;
    constexpr int W = (2 + 3 * C + (C + 3) / 4 + 4 * NL) | 1;
    constexpr int SSLOT = 2 + 3 * C + (C + 3) / 4;
    constexpr int RAw = NL + 2, ROw = C + 1;
    static_assert(C <= 8, "packed per-core counters");
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x, env0 = blockIdx.x * 32;
    const int OW = p.OH >> 1;
    if (lane == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 32u * W * 4u);
        bulk_g2s(sm, p.state + (size_t)env0 * W, 32u * W * 4u, &bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    uint32_t st[W];
#pragma unroll
    for (int k = 0; k < W; ++k) st[k] = sm[lane * W + k];
    __syncwarp();

    uint32_t *ob = sm + (size_t)lane * OW;
    // background: nobody owns, no offers
#pragma unroll
    for (int r = 0; r < N * C + C; ++r) {
        ob[r * RAw] = 0u;
        ob[r * RAw + 1] = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < NL; ++k) ob[r * RAw + 2 + k] = 0xfffefffeu;
    }
    uint32_t cp[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        cp[j] = job_pair(p, st[2 + 3 * j]);
        const int o = core_owner(st[2 + 3 * j]);
        const int row = o > 0 ? (o - 1) * C + j : N * C + j;
        ob[row * RAw] = 0x00010000u;  // [pad, own = 1]
        ob[row * RAw + 1] = cp[j];
    }
    unsigned long long cnt = 0ull;
#pragma unroll
    for (int s = 0; s < NL; ++s) {
        const uint32_t w3 = st[SSLOT + 4 * s + 3];
        const int c = (int)(w3 & 0xffu);
        if (c != 0) {
            const int j = c - 1, r = off_recip(w3);
            const int row = r > 0 ? (r - 1) * C + j : N * C + j;
            const int n = (int)((cnt >> (8 * j)) & 0xffull);
            cnt += 1ull << (8 * j);
            ob[row * RAw + 2 + n] = pair16(off_price(w3), job_rem(st[SSLOT + 4 * s]));
        }
    }
    uint32_t *oo = ob + (N * C + C) * RAw;
#pragma unroll
    for (int s = 0; s < NL; ++s) {
#pragma unroll
        for (int j = 0; j < C; ++j) oo[s * ROw + j] = cp[j];
        oo[s * ROw + C] = job_pair(p, st[SSLOT + 4 * s]);
    }
    // tail padding words of the record are never read; zero them once so the store is clean
    for (int k = (N * C + C) * RAw + NL * ROw; k < OW; ++k) ob[k] = 0u;

    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        bulk_s2g(reinterpret_cast<uint32_t *>(p.obs) + (size_t)env0 * OW, sm, 32u * (uint32_t)OW * 4u);
        bulk_commit();
        bulk_wait_read();
    }
}

// ---- generic path: any domain, one thread per env, straight to global memory ----------------
__global__ void observe_kernel_direct(const __grid_constant__ DevParams p)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.Bpad) return;
    const int N = p.N, C = p.C, NL = p.NL, RAw = NL + 2, ROw = C + 1;
    const uint32_t *st = p.state + (size_t)env * p.W;
    const uint32_t *slot = st + p.sSlot;
    uint32_t *ob = reinterpret_cast<uint32_t *>(p.obs) + (size_t)env * (p.OH >> 1);
    for (int r = 0; r < N * C + C; ++r) {
        ob[r * RAw] = 0u;
        ob[r * RAw + 1] = 0xffffffffu;
        for (int k = 0; k < NL; ++k) ob[r * RAw + 2 + k] = 0xfffefffeu;
    }
    for (int j = 0; j < C; ++j) {
        const int o = core_owner(st[2 + 3 * j]);
        const int row = o > 0 ? (o - 1) * C + j : N * C + j;
        ob[row * RAw] = 0x00010000u;
        ob[row * RAw + 1] = job_pair(p, st[2 + 3 * j]);
    }
    unsigned char cnt[64];
    for (int j = 0; j < C; ++j) cnt[j] = 0;
    for (int s = 0; s < NL; ++s) {
        const uint32_t w3 = slot[4 * s + 3];
        const int c = (int)(w3 & 0xffu);
        if (c == 0) continue;
        const int j = c - 1, r = off_recip(w3);
        const int row = r > 0 ? (r - 1) * C + j : N * C + j;
        ob[row * RAw + 2 + cnt[j]++] = pair16(off_price(w3), job_rem(slot[4 * s]));
    }
    uint32_t *oo = ob + (N * C + C) * RAw;
    for (int s = 0; s < NL; ++s) {
        for (int j = 0; j < C; ++j) oo[s * ROw + j] = job_pair(p, st[2 + 3 * j]);
        oo[s * ROw + C] = job_pair(p, slot[4 * s]);
    }
}

// offer-ID tables: ids [N][C][NL] then auctioneer ids [C][NL] (int16), -2 padding.
// env.correspondingOfferIDs / auctioneer_correspondingOfferIDs, src/SchedulingEnvironment.py:26-29
__global__ void ids_kernel(const __grid_constant__ DevParams p, int16_t *__restrict__ ids)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.B) return;
    const int N = p.N, C = p.C, NL = p.NL;
    const uint32_t *slot = p.state + (size_t)env * p.W + p.sSlot;
    int16_t *o = ids + (size_t)env * (N * C + C) * NL;
    for (int k = 0; k < (N * C + C) * NL; ++k) o[k] = -2;
    unsigned char cnt[64];
    for (int j = 0; j < C; ++j) cnt[j] = 0;
    int id = 0;
    for (int s = 0; s < NL; ++s) {
        const uint32_t w3 = slot[4 * s + 3];
        const int c = (int)(w3 & 0xffu);
        if (c == 0) continue;
        ++id;  // Offer.offerID restarts at 1 every step, src/world.py:324-325
        const int j = c - 1, r = off_recip(w3);
        const int row = r > 0 ? (r - 1) * C + j : N * C + j;
        o[row * NL + cnt[j]++] = (int16_t)id;
    }
}

// Auctioneer.getAuctioneerAction on the current state (src/Auctioneer.py:95-102,
// src/HardcodedModules.py:48-78); same rule and same tie draw as the step kernel's in-kernel auction
__global__ void auctioneer_kernel(const __grid_constant__ DevParams p, int randomTies, int16_t *__restrict__ out)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.B) return;
    const int C = p.C, NL = p.NL;
    const uint32_t *st = p.state + (size_t)env * p.W;
    const uint32_t *slot = st + p.sSlot;
    for (int j = 0; j < C; ++j) {
        int k = NL;
        if (core_owner(st[2 + 3 * j]) == 0) {
            int bn = -1, bd = 1, ncand = 0, first = -1, rank = 0;
            for (int s = 0; s < NL; ++s) {
                const uint32_t w3 = slot[4 * s + 3];
                if ((w3 & 0xffffu) != (uint32_t)(j + 1)) continue;
                int pn = off_price(w3), pd = job_rem(slot[4 * s]);
                if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                const int lhs = pn * bd, rhs = bn * pd;
                if (lhs > rhs) { bn = pn; bd = pd; ncand = 1; first = rank; }
                else if (lhs == rhs) ++ncand;
                ++rank;
            }
            if (first >= 0) {
                k = first;
                if (randomTies && ncand > 1) {
                    uint32_t x[4];
                    env_draw(p, env, kStreamTie, (uint32_t)(j >> 2), 0u, x);  // word j%4 of call j/4
                    const uint32_t xw = (j & 3) == 0 ? x[0] : (j & 3) == 1 ? x[1] : (j & 3) == 2 ? x[2] : x[3];
                int pick = (int)__umulhi(xw, (uint32_t)ncand);
                    rank = 0;
                    for (int s = 0; s < NL && pick >= 0; ++s) {
                        const uint32_t w3 = slot[4 * s + 3];
                        if ((w3 & 0xffffu) != (uint32_t)(j + 1)) continue;
                        int pn = off_price(w3), pd = job_rem(slot[4 * s]);
                        if (pn == -1 || pn == -2 || pd == -1 || pd == -2) { pn = -1; pd = 1; }
                        if (pn * bd == bn * pd) {
                            if (pick == 0) k = rank;
                            --pick;
                        }
                        ++rank;
                    }
                }
            }
        }
        out[(size_t)env * C + j] = (int16_t)k;
    }
}

// reference-shaped dump (debug / parity), see msched_export_state in include/msched.h
struct ExportArgs {
    int env0, count;
    int32_t *core, *slot, *offer, *chain, *chainLen, *misc;
};

__global__ void export_kernel(const __grid_constant__ DevParams p, const ExportArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.count) return;
    const int env = a.env0 + i;
    const int C = p.C, NL = p.NL, K = p.chainCap;
    const uint32_t *st = p.state + (size_t)env * p.W;
    const uint32_t *slot = st + p.sSlot;
    if (a.core)
        for (int j = 0; j < C; ++j) {
            const uint32_t w0 = st[2 + 3 * j];
            const int kind = job_kind(w0);
            int32_t *o = a.core + ((size_t)i * C + j) * 7;
            o[0] = core_owner(w0);
            o[1] = kind >= 0 ? p.prio[kind] : -1;
            o[2] = job_rem(w0);
            o[3] = (int32_t)st[2 + 3 * j + 1];
            o[4] = kind;
            o[5] = (int32_t)st[2 + 3 * j + 2];
            o[6] = kind >= 0 ? p.len[kind] : -1;
        }
    int id = 0;
    for (int s = 0; s < NL; ++s) {
        const uint32_t w0 = slot[4 * s], w3 = slot[4 * s + 3];
        const int kind = job_kind(w0);
        if (a.slot) {
            int32_t *o = a.slot + ((size_t)i * NL + s) * 7;
            o[0] = kind >= 0 ? p.prio[kind] : -1;
            o[1] = job_rem(w0);
            o[2] = (int32_t)slot[4 * s + 1];
            o[3] = kind;
            o[4] = (w3 & 0xffu) != 0u;
            o[5] = (int32_t)slot[4 * s + 2];
            o[6] = kind >= 0 ? p.len[kind] : -1;
        }
        if (a.offer) {
            int32_t *f = a.offer + ((size_t)i * NL + s) * 5;
            if ((w3 & 0xffu) != 0u) {
                f[0] = off_core(w3); f[1] = off_recip(w3); f[2] = off_price(w3); f[3] = job_rem(w0);
                f[4] = ++id;
            } else {
                f[0] = 0; f[1] = f[2] = f[3] = f[4] = -1;
            }
        }
    }
    for (int j = 0; j < C; ++j) {
        const int len = (int)((st[p.sChlen + (j >> 2)] >> ((j & 3) * 8)) & 0xffu);
        if (a.chainLen) a.chainLen[(size_t)i * C + j] = len;
        if (a.chain) {
            const uint32_t *ce = p.chain + ((size_t)env * C + j) * K * 2;
            for (int e = 0; e < K; ++e) {  // newest first like the reference deque
                int32_t *o = a.chain + (((size_t)i * C + j) * K + e) * 5;
                if (e < len) {
                    const int src = len - 1 - e;
                    const uint32_t w1 = ce[2 * src + 1];
                    o[0] = (int)(w1 >> 24);
                    o[1] = src > 0 ? (int)(ce[2 * (src - 1) + 1] >> 24) : 0;
                    o[2] = (int)(int16_t)(w1 & 0xffffu);
                    o[3] = (int)((w1 >> 16) & 0xffu);
                    o[4] = (int32_t)ce[2 * src];
                } else {
                    o[0] = o[1] = o[2] = o[3] = o[4] = -1;
                }
            }
        }
    }
    if (a.misc) {
        int32_t *m = a.misc + (size_t)i * 4;
        m[0] = (int32_t)st[0]; m[1] = (int32_t)st[1]; m[2] = 0; m[3] = 0;
    }
}

// ---- episode aggregates: sums over the environments of one step's result records ----------------
// The reference's train loop accumulates every reward array per step and averages per episode
// (src/trainPPO.py:172-227).  Batched, the per-step arrays are reduced over the env dimension on
// the device and ACCUMULATED into `out` (float64 [result_words + 6]): word k of the record lands
// in out[k] (float fields as float, integer fields as integer; the two quality words hold the sum
// of quality_sum in the first), the packed counts in the six tail entries
//   +0 sum quality_cnt, +1 sum n_accepted, +2 sum n_terminated, +3 sum done,
//   +4 sum over envs with quality_cnt > 0 of the step's mean quality, +5 number of such envs.
__global__ void result_sums_kernel(const __grid_constant__ DevParams p, const uint32_t *__restrict__ result,
                                   double *__restrict__ out)
{
    extern __shared__ double sacc[];  // [RW + 6]
    const int RW = p.RW, nOut = RW + 6;
    for (int k = threadIdx.x; k < nOut; k += blockDim.x) sacc[k] = 0.0;
    __syncthreads();
    const int nFloat = p.rAcc;  // offer (and price) rewards are float32 and come first in the record
    for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < p.B; env += gridDim.x * blockDim.x) {
        const uint32_t *r = result + (size_t)env * RW;
        for (int k = 0; k < p.rQual; ++k) {
            const uint32_t w = r[k];
            if (w != 0u) atomicAdd(&sacc[k], k < nFloat ? (double)__uint_as_float(w) : (double)(int)w);
        }
        const double q = __longlong_as_double(((long long)r[p.rQual + 1] << 32) | (long long)r[p.rQual]);
        const uint32_t cw = r[p.rCounts];
        const int qc = (int)(cw & 0xffu), na = (int)((cw >> 8) & 0xffu), nt = (int)((cw >> 16) & 0xffu);
        if (q != 0.0) atomicAdd(&sacc[p.rQual], q);
        if (qc) { atomicAdd(&sacc[RW + 0], (double)qc); atomicAdd(&sacc[RW + 4], q / (double)qc); atomicAdd(&sacc[RW + 5], 1.0); }
        if (na) atomicAdd(&sacc[RW + 1], (double)na);
        if (nt) atomicAdd(&sacc[RW + 2], (double)nt);
        if ((cw >> 24) & 1u) atomicAdd(&sacc[RW + 3], 1.0);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nOut; k += blockDim.x)
        if (sacc[k] != 0.0) atomicAdd(&out[k], sacc[k]);
}

// column sums of the episode statistics: int32 [B][cols] -> int64 [cols] (accumulated into out)
__global__ void stats_sums_kernel(const int *__restrict__ stats, int rows, int cols, unsigned long long *__restrict__ out)
{
    const int c = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= cols) return;
    long long acc = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x)
        acc += stats[(size_t)r * cols + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0 && acc != 0) atomicAdd(&out[c], (unsigned long long)acc);
}

}  // namespace msched
