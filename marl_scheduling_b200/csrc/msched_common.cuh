// msched_common.cuh -- record layouts, bit packing, Philox and TMA bulk-copy helpers shared by
// the sm_100a kernels of the batched scheduling environment.
//
// State record (uint32 words, one record per environment, env-major so that a tile of envs is
// ONE contiguous chunk that a single cp.async.bulk moves between HBM and shared memory):
//   [0]            next jobID        (Job.IDCounter, src/world.py:80,95-96 -- per world here)
//   [1]            sticky fault flags
//   [2 + 3c + 0]   core c: owner u8 | kind i8 << 8 | remainingLength i16 << 16   (src/world.py:27-37)
//   [2 + 3c + 1]   core c: jobID   (-1 empty)
//   [2 + 3c + 2]   core c: birthDate
//   [S_CHLEN ...]  liability-chain lengths, u8 x 4 per word                     (src/world.py:238)
//   [S_SLOT + 4s + 0] slot s=(agent*L+q): kind i8 | remainingLength i16 << 16   (src/world.py:117-141)
//   [S_SLOT + 4s + 1] jobID ; [+2] birthDate
//   [S_SLOT + 4s + 3] pending offer: coreID u8 (0 = none) | recipientID u8 << 8 | price i16 << 16
//                     (src/world.py:156-196; `wait` == pending offer exists, necessaryTime == the
//                     slot's remainingLength, prio1/jobKind/jobID are the slot's own -- invariant I4)
// priority and initialLength are functions of the kind and live in the kernel parameters.
// The record length is forced ODD so that lane-per-env accesses to the staged tile in shared
// memory (address = lane*W + field) are bank-conflict free for any field index.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msched.h"

namespace msched {

constexpr int kMaxKinds = MSCHED_MAX_KINDS;

struct DevParams {
    // domain
    int B, Bpad, N, C, L, NL, J;
    int newJobs, mult, episodeLength, freePrices, mode, chainCap, auctionMode, spawnMode;
    // layouts
    int W;        // state words per env (odd)
    int AH;       // action halfs per env
    int RW;       // result words per env (odd)
    int OH;       // obs halfs per env
    int sChlen, sSlot;
    int aAcc, aOffc, aOffp, aAuc, aSpawn;
    int rOffer, rPrice, rAcc, rAuc, rAgent, rQual, rCounts, rFlags, rAucIdx, RL, RC;
    // per-kind tables
    int prio[kMaxKinds], len[kMaxKinds], fix[kMaxKinds];
    double cum[kMaxKinds];
    unsigned long long cumThr[kMaxKinds];  // ceil(cum * 2^32): u = x * 2^-32 < cum  <=>  x < cumThr (exact)
    float netZero;
    // dynamic
    int round;
    int doneFlag;  // (round + 1) % episodeLength == 0, src/SchedulingEnvironment.py:64-67
    unsigned long long seed;
    long long envOffset;
    // buffers
    uint32_t *state;
    uint32_t *chain;
    const int16_t *action;
    const double *spawnU;
    uint32_t *result;
    int16_t *obs;
    unsigned long long *timeline;  // diagnostics: 8 x u64 per CTA (smid, clock64 at phase ends), or null
    int *roundDev;                 // device-side round counter (CUDA-graph replays), or null -> `round`
    unsigned *roundTicket;         // CTA exit tickets: the last CTA of a launch advances roundDev (or null)
    int *stats;                    // episode statistics int32 [Bpad][J][4], or null (msched_bind_stats)
    int16_t *cobs;                 // compact observation records (warp kernel / msched_observe_compact), or null
    int COH;                       // compact observation halfs per env
    // compact result record (msched_step_host_compact): int16 / half planes instead of int32 / float32
    uint32_t *cres;                // compact result records, or null -> the full record goes to `result`
    int CW;                        // compact result words per env
    int cOffer, cPrice, cAcc, cAuc, cAgent, cTail;  // half offsets of the planes; word offset of [quality f32, counts, flags]
    // multi-step launches of the fused kernel (msched_step_multi): nSteps consecutive steps per launch, step t reads
    // action + t * actStep (int16), writes result + t * resStep (uint32) and, if obsEvery, obs + t * obsStep (int16);
    // without obsEvery only the last step builds observations and the state tile stays in shared memory in between
    int nSteps, obsEvery;
    long long actStep, resStep, obsStep;
    // ... with the hard-coded agents in the loop (msched_rollout_hardcoded): the actions of step t+1 are computed from
    // the observation tile of step t inside the kernel; the actions after the last step go to actionOut
    int hcPolicy, hcRandomTies, hcOAcc, hcOOff, hcAccRow, hcOffRow;
    int16_t *actionOut;
};

// ---- host+device layout arithmetic ------------------------------------------------------
__host__ __device__ inline int make_odd(int x) { return (x & 1) ? x : x + 1; }

// ---- bit packing --------------------------------------------------------------------------
__device__ __forceinline__ int core_owner(uint32_t w) { return (int)(w & 0xffu); }
__device__ __forceinline__ int job_kind(uint32_t w) { return (int)(int8_t)((w >> 8) & 0xffu); }
__device__ __forceinline__ int job_rem(uint32_t w) { return (int)(int16_t)(w >> 16); }
__device__ __forceinline__ uint32_t pack_core(int owner, int kind, int rem)
{
    return (uint32_t)(owner & 0xff) | ((uint32_t)(kind & 0xff) << 8) | ((uint32_t)(rem & 0xffff) << 16);
}
// slot word 0 uses the same kind/rem positions with the owner byte left zero
__device__ __forceinline__ uint32_t pack_slot(int kind, int rem) { return pack_core(0, kind, rem); }
constexpr uint32_t kEmptyJobW0 = 0xffffff00u;  // kind -1, rem -1, owner 0
constexpr uint32_t kEmptyId = 0xffffffffu;     // jobID / birthDate -1

__device__ __forceinline__ int off_core(uint32_t w) { return (int)(w & 0xffu); }
__device__ __forceinline__ int off_recip(uint32_t w) { return (int)((w >> 8) & 0xffu); }
__device__ __forceinline__ int off_price(uint32_t w) { return (int)(int16_t)(w >> 16); }
__device__ __forceinline__ uint32_t pack_offer(int core, int recip, int price)
{
    return (uint32_t)(core & 0xff) | ((uint32_t)(recip & 0xff) << 8) | ((uint32_t)(price & 0xffff) << 16);
}
// chain entry: word0 = round, word1 = price i16 | necessaryTime u8 << 16 | offererID u8 << 24;
// the recipient is the previous (older) entry's offerer, 0 for the oldest (src/world.py:261-293)
__device__ __forceinline__ uint32_t pack_chain(int price, int time, int offerer)
{
    return (uint32_t)(price & 0xffff) | ((uint32_t)(time & 0xff) << 16) | ((uint32_t)(offerer & 0xff) << 24);
}

// ---- Philox4x32-10 (Salmon et al., SC'11) ------------------------------------------------
// Device randomness contract: counter = (global env lo, global env hi, round,
// stream<<28 | a<<12 | b), key = seed.  Identical, independently written, in oracle/.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t kStreamSpawn = 0, kStreamTie = 1, kStreamPolicy = 2;

// world.round of this step: a kernel argument, or (graph-replay mode) a device counter that a
// one-thread kernel advances after every step
__device__ __forceinline__ int cur_round(const DevParams &p) { return p.roundDev ? *p.roundDev : p.round; }
__device__ __forceinline__ int cur_done(const DevParams &p, int round)
{   // done = (world.round % episodeLength == 0) after the step, src/SchedulingEnvironment.py:64-67
    return p.roundDev ? (((round + 1) % p.episodeLength) == 0 ? 1 : 0) : p.doneFlag;
}

__device__ __forceinline__ void env_draw_at(const DevParams &p, int round, int env, uint32_t stream, uint32_t a,
                                            uint32_t b, uint32_t out[4])
{
    const unsigned long long g = (unsigned long long)(p.envOffset + env);
    philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)round, (stream << 28) | (a << 12) | b,
                  (uint32_t)p.seed, (uint32_t)(p.seed >> 32), out);
}
__device__ __forceinline__ void env_draw(const DevParams &p, int env, uint32_t stream, uint32_t a,
                                         uint32_t b, uint32_t out[4])
{
    env_draw_at(p, cur_round(p), env, stream, a, b, out);
}
__device__ __forceinline__ double u53(const uint32_t x[4])
{
    const unsigned long long v = ((unsigned long long)x[1] << 32) | x[0];
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) + mbarrier --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MSCHED_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MSCHED_DONE_%=;\n"
        "bra MSCHED_WAIT_%=;\n"
        "MSCHED_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion signalled on the mbarrier (bytes % 16 == 0, both 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// wait only until the bulk stores have READ their shared-memory source (enough before CTA exit)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned smid()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long r;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r));
    return r;
}
// Called by one thread of every CTA after its last use of the round: the CTA that draws the last
// ticket knows that every other CTA of the launch is past its reads and advances the counter, so a
// step needs no separate "round += 1" launch (graph-replay mode only)
// The same in two halves for a kernel whose threads ALL hold the round in a register before a CTA barrier: the
// ticket is drawn right after that barrier (its latency hides behind the step) and only redeemed at the end
__device__ __forceinline__ unsigned take_round_ticket(const DevParams &p)
{
    if (!p.roundTicket) return 0u;
    __threadfence();
    return atomicAdd(p.roundTicket, 1u);
}
__device__ __forceinline__ void redeem_round_ticket(const DevParams &p, unsigned t, int steps = 1)
{
    if (p.roundTicket && t == gridDim.x - 1) {  // every CTA of the launch has read the round
        *p.roundTicket = 0u;
        *p.roundDev += steps;
        __threadfence();
    }
}
__device__ __forceinline__ void finish_round(const DevParams &p)
{
    if (p.roundTicket) {
        __threadfence();
        const unsigned t = atomicAdd(p.roundTicket, 1u);
        if (t == gridDim.x - 1) {
            *p.roundTicket = 0u;
            *p.roundDev += 1;
            __threadfence();
        }
    }
}

// Episode statistics the train scripts derive from world.acceptedOffers / world.verweilzeiten
// (src/trainPPO.py:172-227, src/world.py:350-357): per env and job kind
// [sum of accepted prices, #accepted, sum of (dwell - 1), #terminated].  Rare events (about one per
// env-step): fire-and-forget reductions straight to HBM.
__device__ __forceinline__ void stat_accept(const DevParams &p, int env, int kind, int price)
{
    if (p.stats && kind >= 0) {
        int *sp = p.stats + ((size_t)env * p.J + kind) * 4;
        atomicAdd(sp, price);
        atomicAdd(sp + 1, 1);
    }
}
__device__ __forceinline__ void stat_terminate(const DevParams &p, int env, int kind, int round, uint32_t birth)
{
    if (p.stats && kind >= 0) {
        int *sp = p.stats + ((size_t)env * p.J + kind) * 4;
        atomicAdd(sp + 2, round - (int)birth - 1);
        atomicAdd(sp + 3, 1);
    }
}
// make this thread's generic-proxy shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace msched
