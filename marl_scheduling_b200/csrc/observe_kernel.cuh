// observe_kernel.cuh -- dense reference-layout observations and the reference-shaped state dump.
//
//   src/Agent.py:148-212    DividedAgent.gatherObservations / getAcceptorObservationTensorAndIDs
//   src/Agent.py:271-300    getOfferNetObservationTensor
//   src/Auctioneer.py:20-77 gatherDividedAuctioneerObservation
//
// Observation record (int16 per env, layout from msched_get_layout):
//   acceptor [N][C][3+2NL] : own?1:0, own?prio:-1, own?rem:-1, (price,time) of the offers addressed
//                            to (agent, core) in creation order, (-2,-2) padding
//   offer    [N][L][2C+2]  : (prio, rem) of every core, then of the slot (empty = -1)
//   auctioneer [C][3+2NL]  : as acceptor with owner 0
//   ids      [N][C][NL], auctioneer ids [C][NL] : offerIDs of those offers, -2 padding
// The semi-/fully-aggregated layouts (src/Agent.py:82-140, 399-461) are concatenations of these
// blocks and are assembled as views on the host side.
#pragma once
#include "msched_common.cuh"

namespace msched {

// one acceptor-style row for `who` (agentID, or 0 = auctioneer) and core j
__device__ __forceinline__ void acceptor_row(const DevParams &p, const uint32_t *st, int who, int j,
                                             int16_t *row, int16_t *ids)
{
    const int NL = p.NL;
    const uint32_t *slot = st + p.sSlot;
    const uint32_t cw0 = st[2 + 3 * j];
    const bool own = core_owner(cw0) == who;
    const int kind = job_kind(cw0);
    row[0] = own ? 1 : 0;
    row[1] = (int16_t)((own && kind >= 0) ? p.prio[kind] : -1);
    row[2] = (int16_t)(own ? job_rem(cw0) : -1);
    int n = 0;
    if (own) {  // every pending offer to core j is addressed to its owner
        const uint32_t key = (uint32_t)(j + 1) | ((uint32_t)who << 8);
        int id = 0;
        for (int s = 0; s < NL; ++s) {
            const uint32_t w3 = slot[4 * s + 3];
            if ((w3 & 0xffu) == 0u) continue;
            ++id;  // Offer.offerID restarts at 1 every step, src/world.py:324-325
            if ((w3 & 0xffffu) != key) continue;
            row[3 + 2 * n] = (int16_t)off_price(w3);
            row[4 + 2 * n] = (int16_t)job_rem(slot[4 * s]);
            ids[n] = (int16_t)id;
            ++n;
        }
    }
    for (; n < NL; ++n) {
        row[3 + 2 * n] = -2;
        row[4 + 2 * n] = -2;
        ids[n] = -2;
    }
}

__device__ __forceinline__ void observe_env(const DevParams &p, const uint32_t *st, int16_t *ob)
{
    const int N = p.N, C = p.C, L = p.L, NL = p.NL, Wd = 3 + 2 * NL, Wo = 2 * C + 2;
    const uint32_t *slot = st + p.sSlot;
    for (int a = 0; a < N; ++a)
        for (int j = 0; j < C; ++j)
            acceptor_row(p, st, a + 1, j, ob + p.oAcc + (a * C + j) * Wd, ob + p.oIds + (a * C + j) * NL);
    for (int j = 0; j < C; ++j)
        acceptor_row(p, st, 0, j, ob + p.oAuc + j * Wd, ob + p.oAucIds + j * NL);
    for (int s = 0; s < NL; ++s) {
        int16_t *row = ob + p.oOff + s * Wo;
        for (int j = 0; j < C; ++j) {
            const uint32_t cw0 = st[2 + 3 * j];
            const int kind = job_kind(cw0);
            row[2 * j] = (int16_t)(kind >= 0 ? p.prio[kind] : -1);
            row[2 * j + 1] = (int16_t)job_rem(cw0);
        }
        const uint32_t w0 = slot[4 * s];
        const int kind = job_kind(w0);
        row[2 * C] = (int16_t)(kind >= 0 ? p.prio[kind] : -1);
        row[2 * C + 1] = (int16_t)job_rem(w0);
    }
}

// staged variant: state tile in by TMA, observation tile out by TMA.
// dynamic smem = blockDim.x * (W*4 + OH*2)
__global__ void __launch_bounds__(128) observe_kernel_staged(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int T = blockDim.x, lane = threadIdx.x, env0 = blockIdx.x * T;
    const uint32_t stBytes = (uint32_t)T * p.W * 4u, obBytes = (uint32_t)T * p.OH * 2u;
    uint32_t *sState = reinterpret_cast<uint32_t *>(smem);
    int16_t *sObs = reinterpret_cast<int16_t *>(smem + stBytes);
    if (lane == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, stBytes);
        bulk_g2s(sState, p.state + (size_t)env0 * p.W, stBytes, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    observe_env(p, sState + (size_t)lane * p.W, sObs + (size_t)lane * p.OH);
    fence_async_smem();
    __syncthreads();
    if (lane == 0) {
        bulk_s2g(p.obs + (size_t)env0 * p.OH, sObs, obBytes);
        bulk_commit();
        bulk_wait_all();
    }
}

// direct variant for domains whose observation tile does not fit in shared memory
__global__ void observe_kernel_direct(const __grid_constant__ DevParams p)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.Bpad) return;
    observe_env(p, p.state + (size_t)env * p.W, p.obs + (size_t)env * p.OH);
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

}  // namespace msched
