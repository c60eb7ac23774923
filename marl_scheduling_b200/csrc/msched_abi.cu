// msched_abi.cu -- the C-ABI (include/msched.h) over the sm_100a kernels.
// There is no CPU fallback anywhere in this library: without a CUDA device every launch entry
// point returns MSCHED_E_NODEVICE.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "abi_common.h"
#include "msched_common.cuh"
#include "observe_kernel.cuh"
#include "step_kernel.cuh"
#include "coop_step_kernel.cuh"
#include "fused_step_kernel.cuh"
#include "hardcoded_kernel.cuh"
#include "warp_step.h"

using namespace msched;

thread_local std::string msched_g_err;

namespace {

int even_odd_half(int halfs)
{  // even count of int16 whose word count is odd (bank-conflict-free lane stride)
    if (halfs & 1) ++halfs;
    if (((halfs / 2) & 1) == 0) halfs += 2;
    return halfs;
}

int compute_layout(const MschedConfig *c, MschedLayout *o)
{
    if (!c || !o) return fail(MSCHED_E_ARG, "null config/layout");
    if (c->abi_version != MSCHED_ABI_VERSION) return fail(MSCHED_E_ARG, "abi_version mismatch");
    if (c->B < 1) return fail(MSCHED_E_ARG, "B must be >= 1");
    if (c->N < 1 || c->N > 250) return fail(MSCHED_E_ARG, "numberOfAgents must be in 1..250");
    if (c->C < 1 || c->C > 64) return fail(MSCHED_E_ARG, "numberOfCores must be in 1..64");
    if (c->L < 1 || (long long)c->N * c->L > 4096) return fail(MSCHED_E_ARG, "N*L must be in 1..4096");
    if (c->J < 1 || c->J > MSCHED_MAX_KINDS) return fail(MSCHED_E_ARG, "job kinds must be in 1..16");
    if (c->newJobsPerRound < 0 || c->newJobsPerRound > c->L)
        return fail(MSCHED_E_ARG, "newJobsPerRoundPerAgent must be in 0..collectionLength");
    if (c->episodeLength < 1) return fail(MSCHED_E_ARG, "episodeLength must be >= 1");
    if (c->chainCapacity < 1 || c->chainCapacity > 255)
        return fail(MSCHED_E_ARG, "chainCapacity must be in 1..255");
    if (c->rewardVariant < 0 || c->rewardVariant > 3) return fail(MSCHED_E_ARG, "bad rewardVariant");
    if (c->auctionMode < 0 || c->auctionMode > 2) return fail(MSCHED_E_ARG, "bad auctionMode");
    if (c->spawnMode < 0 || c->spawnMode > 2) return fail(MSCHED_E_ARG, "bad spawnMode");
    const bool freeVar = c->rewardVariant == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL ||
                         c->rewardVariant == MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL;
    if (freeVar != (c->freePrices != 0))
        return fail(MSCHED_E_ARG, "freePrices must match the reward variant");
    for (int k = 0; k < c->J; ++k) {
        if (c->len[k] < 1 || c->len[k] > 255) return fail(MSCHED_E_ARG, "job lengths must be in 1..255");
        if (c->prio[k] < -32000 || c->prio[k] > 32000) return fail(MSCHED_E_ARG, "priority out of int16 range");
        if (c->fixPrice[k] < -32000 || c->fixPrice[k] > 32000) return fail(MSCHED_E_ARG, "price out of int16 range");
    }
    const int N = c->N, C = c->C, L = c->L, NL = N * L;
    const bool agg = c->rewardVariant == MSCHED_REWARD_AGGREGATED_FIXED;
    memset(o, 0, sizeof(*o));
    o->padded_envs = msched_padded_envs(c->B);
    o->state_words = make_odd(2 + 3 * C + (C + 3) / 4 + 4 * NL);
    int h = 0;
    o->a_acceptor = h; h += N * C;
    o->a_offer_core = h; h += NL;
    o->a_offer_price = -1;
    if (c->freePrices) { o->a_offer_price = h; h += NL; }
    o->a_auctioneer = -1;
    if (c->auctionMode == MSCHED_AUCTION_EXTERNAL) { o->a_auctioneer = h; h += C; }
    o->a_spawn_kind = -1;
    if (c->spawnMode == MSCHED_SPAWN_KINDS) { o->a_spawn_kind = h; h += N * c->newJobsPerRound; }
    o->action_halfs = even_odd_half(h);
    o->RL = agg ? 1 : L;
    o->RC = agg ? 1 : C;
    int w = 0;
    o->r_offer = w; w += N * o->RL;
    o->r_price = -1;
    if (c->freePrices) { o->r_price = w; w += N * o->RL; }
    o->r_acceptor = w; w += N * o->RC;
    o->r_auctioneer = w; w += C;
    o->r_agent = w; w += N;
    o->r_quality = w; w += 2;
    o->r_counts = w; w += 1;
    o->r_flags = w; w += 1;
    o->r_auctioneer_idx = w; w += (C + 1) / 2;
    o->result_words = make_odd(w);
    const int RA = 3 + 2 * NL + 1, RO = 2 * C + 2;
    const long long oh = (long long)(N * C + C) * RA + (long long)NL * RO;
    if (oh > (1ll << 28)) return fail(MSCHED_E_ARG, "observation record too large");
    o->o_acceptor = 1;
    o->o_auctioneer = N * C * RA + 1;
    o->o_offer = (N * C + C) * RA;
    o->o_acc_row = RA;
    o->o_off_row = RO;
    o->obs_halfs = even_odd_half((int)oh);
    o->ids_halfs = (N * C + C) * NL;
    o->cobs_halfs = compact_obs_halfs(C, NL);
    o->c_core = 0;
    o->c_slot = 4 * C;
    o->c_offer = 4 * C + 2 * NL;
    o->chain_words = C * c->chainCapacity * 2;
    return MSCHED_OK;
}

int compute_compact_layout(const MschedConfig *c, const MschedLayout &l, MschedCompactResultLayout *o)
{
    if (!o) return fail(MSCHED_E_ARG, "null layout");
    for (int k = 0; k < c->J; ++k)
        if (c->prio[k] < -1024 || c->prio[k] > 1024) return fail(MSCHED_E_ARG, "compact results: priorities beyond the exact half range");
    {
        const float z = (float)c->netZeroOfferReward;
        if (__half2float(__float2half_rn(z)) != z || (double)z != c->netZeroOfferReward)
            return fail(MSCHED_E_ARG, "compact results: netZeroOfferReward is not representable as IEEE half");
    }
    int h = 0;
    o->c_offer = h; h += c->N * l.RL;
    o->c_price = -1;
    if (c->freePrices) { o->c_price = h; h += c->N * l.RL; }
    o->c_acceptor = h; h += c->N * l.RC;
    o->c_auctioneer = h; h += c->C;
    o->c_agent = h; h += c->N;
    const int w = (h + 1) / 2;
    // tail: quality_sum (float32), then ONE word: counts in bits 0..24 as in the full record, the sticky flags in bits
    // 25..31.  The word count is not forced odd: config 3's 14 words make a 32-env tile 1,792 bytes = 7 x 256, so every
    // tile the kernel writes over PCIe starts on a 256-byte boundary (the 2-way bank conflict of an even stride costs
    // a few dozen shared-memory stores per tile)
    o->c_quality = w; o->c_counts = w + 1; o->c_flags = w + 1;
    o->words = w + 2;
    return MSCHED_OK;
}

typedef void (*StepKernel)(const DevParams);
typedef void (*ObsKernel)(const DevParams);

struct Handle {
    MschedConfig cfg;
    MschedLayout lay;
    DevParams p;
    int device;
    long long round;
    int smemOptin;
    int stepTile;  // envs per CTA of the lane-per-env step kernel (0 = not available)
    StepKernel stepFn;
    StepKernel coopFn;  // cooperative kernel (G lanes per env), null = not available
    int coopG, coopThreads;
    size_t coopSmem;
    bool useCoop;
    MschedCompactResultLayout clay;  // compact result record (words == 0: not available for this configuration)
    size_t warpSmem;  // warp-per-environment kernel (msched_warp.cu): shared memory per CTA, 0 = not available
    bool useWarp;
    StepKernel fusedFn;  // compile-time-domain register-resident kernel (step + observations), or null
    StepKernel multiFn = nullptr;  // its multi-step instantiation (msched_step_multi), or null
    StepKernel hcFn = nullptr;     // ... with the hard-coded agents in the loop (msched_rollout_hardcoded), or null
    size_t fusedSmem, fusedSmemObs;  // dynamic shared memory without / with the observation tile (compact result tile included)
    size_t hostSmem;     // dynamic shared memory of a host-buffer launch (holds its resident CTAs per SM down), 0 = as computed
    size_t fusedSmemNC, fusedSmemObsNC;  // the same for launches without the compact result tile (p.cres null): cfg3 13.6 KB
                                         // with the reserve instead of 15.2 KB, i.e. 16 resident CTAs per SM instead of 15
    int fusedRoles;                  // warps per 32-env tile
    bool useFused, fuseObs;
    ObsKernel obsFn;  // compile-time-domain observation kernel, or null -> direct kernel
    int16_t *stageAction;
    uint32_t *stageResult;
    int *roundDev;       // device-side round counter + CTA ticket word (msched_set_round_mode)
    bool deviceRound;
    cudaStream_t hostStream[2];  // msched_step_host pipelines its chunks over these
    cudaEvent_t evStart, evDone[2];
};

void free_handle(Handle *h)
{
    if (!h) return;
    cudaFree(h->stageAction);
    cudaFree(h->stageResult);
    cudaFree(h->roundDev);
    for (int k = 0; k < 2; ++k) {
        if (h->hostStream[k]) cudaStreamDestroy(h->hostStream[k]);
        if (h->evDone[k]) cudaEventDestroy(h->evDone[k]);
    }
    if (h->evStart) cudaEventDestroy(h->evStart);
    delete h;
}
// msched_create returns early on any CUDA error: the guard releases what was acquired so far
struct HandleGuard {
    Handle *h;
    ~HandleGuard() { free_handle(h); }
    Handle *release() { Handle *r = h; h = nullptr; return r; }
};

StepKernel pick_step_kernel(int N, int C, int L)
{
    if (N == 2 && C == 3 && L == 3) return step_kernel<2, 3, 3>;  // BASELINE cfg3 (Exp 4-2)
    if (N == 4 && C == 4 && L == 3) return step_kernel<4, 4, 3>;  // BASELINE cfg2 / cfg4
    if (N == 2 && C == 3 && L == 2) return step_kernel<2, 3, 2>;  // BASELINE cfg1 (trainHC)
    if (N == 2 && C == 2 && L == 3) return step_kernel<2, 2, 3>;  // README 2-agent domain
    return step_kernel<0, 0, 0>;
}

template <int R>
StepKernel pick_fused_kernel_r(int N, int C, int L)
{
    if (N == 2 && C == 3 && L == 3) return fused_step_kernel<2, 3, 3, R>;  // BASELINE cfg3 (Exp 4-2)
    if (N == 4 && C == 4 && L == 3) return fused_step_kernel<4, 4, 3, R>;  // BASELINE cfg2 / cfg4
    if (N == 2 && C == 3 && L == 2) return fused_step_kernel<2, 3, 2, R>;  // BASELINE cfg1 (trainHC)
    if (N == 2 && C == 2 && L == 3) return fused_step_kernel<2, 2, 3, R>;  // README 2-agent domain
    return nullptr;
}

// compile-time configurations of the BASELINE runs (device Philox spawn, one new job per agent and round, in-kernel
// auction with random ties): cfg3 = free prices + commercial reward, cfg2 / cfg4 = fixed prices
StepKernel pick_fused_spec(const MschedConfig &c, int roles)
{
    if (c.spawnMode != MSCHED_SPAWN_PHILOX || c.newJobsPerRound != 1 || c.auctionMode != MSCHED_AUCTION_RANDOM_MAX) return nullptr;
    if (c.N == 2 && c.C == 3 && c.L == 3 && c.rewardVariant == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL) {
        constexpr int S = fused_spec(MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL, MSCHED_AUCTION_RANDOM_MAX, MSCHED_SPAWN_PHILOX, 1);
        return roles == 1 ? fused_step_kernel<2, 3, 3, 1, S> : roles == 2 ? fused_step_kernel<2, 3, 3, 2, S> : fused_step_kernel<2, 3, 3, 4, S>;
    }
    if (c.N == 4 && c.C == 4 && c.L == 3 && c.rewardVariant == MSCHED_REWARD_DIVIDED_FIXED) {
        constexpr int S = fused_spec(MSCHED_REWARD_DIVIDED_FIXED, MSCHED_AUCTION_RANDOM_MAX, MSCHED_SPAWN_PHILOX, 1);
        return roles == 2 ? fused_step_kernel<4, 4, 3, 2, S> : roles == 4 ? fused_step_kernel<4, 4, 3, 4, S> : nullptr;
    }
    return nullptr;
}

// the multi-step instantiations (msched_step_multi): the BASELINE configurations and the generic 2-role kernels of
// their domains
template <bool HC>
StepKernel pick_fused_multi_t(const MschedConfig &c, int roles, bool spec)
{
    const bool baseline = spec && c.spawnMode == MSCHED_SPAWN_PHILOX && c.newJobsPerRound == 1 && c.auctionMode == MSCHED_AUCTION_RANDOM_MAX;
    if (c.N == 2 && c.C == 3 && c.L == 3 && (roles == 2 || roles == 4)) {
        constexpr int S = fused_spec(MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL, MSCHED_AUCTION_RANDOM_MAX, MSCHED_SPAWN_PHILOX, 1);
        if constexpr (!HC)  // (the hard-coded agents play fixed prices)
            if (baseline && c.rewardVariant == MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL)
                return roles == 2 ? fused_step_kernel<2, 3, 3, 2, S, true> : fused_step_kernel<2, 3, 3, 4, S, true>;
        return roles == 2 ? fused_step_kernel<2, 3, 3, 2, -1, true, HC> : fused_step_kernel<2, 3, 3, 4, -1, true, HC>;
    }
    if (c.N == 4 && c.C == 4 && c.L == 3 && roles == 4) {
        constexpr int S = fused_spec(MSCHED_REWARD_DIVIDED_FIXED, MSCHED_AUCTION_RANDOM_MAX, MSCHED_SPAWN_PHILOX, 1);
        if (baseline && c.rewardVariant == MSCHED_REWARD_DIVIDED_FIXED) return fused_step_kernel<4, 4, 3, 4, S, true, HC>;
        return fused_step_kernel<4, 4, 3, 4, -1, true, HC>;
    }
    if (c.N == 2 && c.C == 3 && c.L == 2 && (roles == 2 || roles == 4))
        return roles == 2 ? fused_step_kernel<2, 3, 2, 2, -1, true, HC> : fused_step_kernel<2, 3, 2, 4, -1, true, HC>;
    return nullptr;
}
StepKernel pick_fused_multi(const MschedConfig &c, int roles, bool spec) { return pick_fused_multi_t<false>(c, roles, spec); }
StepKernel pick_fused_hc(const MschedConfig &c, int roles, bool spec) { return pick_fused_multi_t<true>(c, roles, spec); }

StepKernel pick_fused_kernel(int N, int C, int L, int roles)
{
    return roles == 1 ? pick_fused_kernel_r<1>(N, C, L) : roles == 2 ? pick_fused_kernel_r<2>(N, C, L)
                                                                     : pick_fused_kernel_r<4>(N, C, L);
}

ObsKernel pick_obs_kernel(int N, int C, int L)
{
    if (N == 2 && C == 3 && L == 3) return observe_kernel_t<2, 3, 3>;
    if (N == 4 && C == 4 && L == 3) return observe_kernel_t<4, 4, 3>;
    if (N == 2 && C == 3 && L == 2) return observe_kernel_t<2, 3, 2>;
    if (N == 2 && C == 2 && L == 3) return observe_kernel_t<2, 2, 3>;
    return nullptr;
}

void fill_params(Handle *h)
{
    const MschedConfig &c = h->cfg;
    const MschedLayout &l = h->lay;
    DevParams &p = h->p;
    memset(&p, 0, sizeof(p));
    p.B = c.B; p.Bpad = l.padded_envs; p.N = c.N; p.C = c.C; p.L = c.L; p.NL = c.N * c.L; p.J = c.J;
    p.newJobs = c.newJobsPerRound; p.mult = c.rewardMultiplier; p.episodeLength = c.episodeLength;
    p.freePrices = c.freePrices; p.mode = c.rewardVariant; p.chainCap = c.chainCapacity;
    p.auctionMode = c.auctionMode; p.spawnMode = c.spawnMode;
    p.W = l.state_words; p.AH = l.action_halfs; p.RW = l.result_words; p.OH = l.obs_halfs;
    p.sChlen = 2 + 3 * c.C; p.sSlot = p.sChlen + (c.C + 3) / 4;
    p.aAcc = l.a_acceptor; p.aOffc = l.a_offer_core; p.aOffp = l.a_offer_price; p.aAuc = l.a_auctioneer;
    p.aSpawn = l.a_spawn_kind;
    p.rOffer = l.r_offer; p.rPrice = l.r_price; p.rAcc = l.r_acceptor; p.rAuc = l.r_auctioneer;
    p.rAgent = l.r_agent; p.rQual = l.r_quality; p.rCounts = l.r_counts; p.rFlags = l.r_flags;
    p.rAucIdx = l.r_auctioneer_idx; p.RL = l.RL; p.RC = l.RC;
    for (int k = 0; k < MSCHED_MAX_KINDS; ++k) {
        p.prio[k] = c.prio[k]; p.len[k] = c.len[k]; p.fix[k] = c.fixPrice[k]; p.cum[k] = c.cumProb[k];
        {
            double t = ceil(c.cumProb[k] * 4294967296.0);
            if (!(t > 0.0)) t = 0.0;
            if (t > 4294967296.0) t = 4294967296.0;
            p.cumThr[k] = (unsigned long long)t;
        }
    }
    p.netZero = (float)c.netZeroOfferReward;
    p.seed = c.seed;
    p.envOffset = c.envOffset;
    p.COH = l.cobs_halfs;
}

int pick_tile(size_t bytesPerEnv, int smemOptin, const char *envName)
{
    if (const char *e = getenv(envName)) {
        const int t = atoi(e);
        if ((t == 32 || t == 64 || t == 128) && (size_t)t * bytesPerEnv + 64 <= (size_t)smemOptin) return t;
    }
    // prefer 64-env tiles with >= 2 CTAs per SM, then 32, then 128 never needed
    if (64 * bytesPerEnv <= 100 * 1024) return 64;
    if (32 * bytesPerEnv + 64 <= (size_t)smemOptin) return 32;
    return 0;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// launch the step kernel for the env range described by p (p.Bpad padded envs starting at p.state)
// selfAdvance: the launch covers the whole step, so its last CTA advances the device-side round
// hostIO: the action and / or result tiles of this launch live in pinned HOST memory (zero-copy): the launch is held
// to a few resident CTAs per SM, so that it runs in waves and the PCIe reads of later tiles overlap the writes of
// earlier ones (the bus is full duplex) -- all 2,048 tiles of a 65,536-environment step resident at once read
// together and then write together (measured, config 3: 1.03e9 agent-steps/s unlimited, 1.12e9 at 9 CTAs per SM,
// 1.19e9 at 5, 1.22e9 at 3 or 2)
void launch_step(const Handle *h, const DevParams &p0, cudaStream_t s, bool selfAdvance = true, bool hostIO = false)
{
    DevParams p = p0;
    p.roundDev = h->deviceRound ? h->roundDev : nullptr;
    p.roundTicket = (h->deviceRound && selfAdvance) ? reinterpret_cast<unsigned *>(h->roundDev + 1) : nullptr;
    if (h->useFused) {
        const size_t smem = p.cres ? (p.obs ? h->fusedSmemObs : h->fusedSmem) : (p.obs ? h->fusedSmemObsNC : h->fusedSmemNC);
        h->fusedFn<<<p.Bpad / 32, 32 * h->fusedRoles, (hostIO && h->hostSmem > smem) ? h->hostSmem : smem, s>>>(p);
    } else if (h->useWarp) {
        launch_warp_step(p, h->warpSmem, s);
    } else if (h->useCoop) {
        const int E = h->coopThreads / h->coopG;
        h->coopFn<<<p.Bpad / E, h->coopThreads, h->coopSmem, s>>>(p);
    } else {
        const int T = h->stepTile;
        const size_t smem = (size_t)T * ((size_t)p.W * 4 + (size_t)p.AH * 2 + (size_t)p.RW * 4 + (size_t)scratch_words(p.C) * 4);
        h->stepFn<<<p.Bpad / T, T, smem, s>>>(p);
    }
}

}  // namespace

extern "C" {

int msched_abi_version(void) { return MSCHED_ABI_VERSION; }
const char *msched_last_error(void) { return msched_g_err.c_str(); }
int msched_padded_envs(int B)
{
    return ((B + MSCHED_TILE_ENVS - 1) / MSCHED_TILE_ENVS) * MSCHED_TILE_ENVS;
}
int msched_get_layout(const MschedConfig *cfg, MschedLayout *out) { return compute_layout(cfg, out); }

int msched_create(const MschedConfig *cfg, int device, void **handle)
{
    if (!handle) return fail(MSCHED_E_ARG, "null handle pointer");
    *handle = nullptr;
    MschedLayout lay;
    int rc = compute_layout(cfg, &lay);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MSCHED_E_NODEVICE, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(MSCHED_E_ARG, "bad device index");
    CUDA_TRY(cudaSetDevice(device));
    {
        double rcp[256];
        unsigned char odd[256];
        for (int t = 0; t < 256; ++t) {
            rcp[t] = t ? 1.0 / (double)t : 0.0;
            int o = t ? t : 1;
            while ((o & 1) == 0) o >>= 1;
            odd[t] = (unsigned char)o;
        }
        CUDA_TRY(cudaMemcpyToSymbol(c_rcp, rcp, sizeof(rcp)));
        CUDA_TRY(cudaMemcpyToSymbol(c_oddpart, odd, sizeof(odd)));
    }
    HandleGuard guard{new Handle()};
    Handle *h = guard.h;
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->lay = lay;
    h->device = device;
    fill_params(h);
    CUDA_TRY(cudaDeviceGetAttribute(&h->smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t stepBytes = (size_t)lay.state_words * 4 + (size_t)lay.action_halfs * 2 + (size_t)lay.result_words * 4 +
                             (size_t)scratch_words(cfg->C) * 4;
    // lane-per-env kernel (needs N*L <= 254 and 32 envs of records in shared memory)
    h->stepTile = (cfg->N * cfg->L <= 254) ? pick_tile(stepBytes, h->smemOptin, "MSCHED_STEP_TILE") : 0;
    if (h->stepTile) {
        h->stepFn = pick_step_kernel(cfg->N, cfg->C, cfg->L);
        CUDA_TRY(cudaFuncSetAttribute(h->stepFn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(h->stepTile * stepBytes)));
    }
    // cooperative kernel: G lanes per env, G = next power of two >= max(C, N), clamped to 4..32
    {
        int G = 4;
        while (G < 32 && G < (cfg->C > cfg->N ? cfg->C : cfg->N)) G <<= 1;
        if (const char *e = getenv("MSCHED_COOP_G")) {
            const int g = atoi(e);
            if (g == 4 || g == 8 || g == 16 || g == 32) G = g;
        }
        const size_t coopBytes = (size_t)lay.state_words * 4 + (size_t)lay.action_halfs * 2 +
                                 (size_t)lay.result_words * 4 + (size_t)coop_scratch_words(cfg->C) * 4;
        int T = 128;
        while (T > 4 * G && (size_t)(T / G) * coopBytes + 64 > (size_t)h->smemOptin) T >>= 1;
        if ((size_t)(T / G) * coopBytes + 64 <= (size_t)h->smemOptin) {
            h->coopG = G;
            h->coopThreads = T;
            h->coopSmem = (size_t)(T / G) * coopBytes;
            h->coopFn = G == 4 ? coop_step_kernel<4> : G == 8 ? coop_step_kernel<8>
                      : G == 16 ? coop_step_kernel<16> : coop_step_kernel<32>;
            CUDA_TRY(cudaFuncSetAttribute(h->coopFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->coopSmem));
        }
    }
    // warp-per-environment kernel: the large domains (BASELINE config 5).  Default for more than 8 cores or more
    // than 64 job slots per environment, where the lane / group kernels' serial offer sweep dominates
    {
        CUDA_TRY(warp_step_init());
        h->warpSmem = warp_step_smem_bytes(h->p, h->smemOptin);
        if (h->warpSmem) CUDA_TRY(warp_step_prepare(h->warpSmem));
        h->useWarp = h->warpSmem && (cfg->C > 8 || cfg->N * cfg->L > 64);
    }
    if (!h->stepTile && !h->coopFn && !h->warpSmem) {
        return fail(MSCHED_E_ARG, "domain too large: the records of 4 environments must fit in shared memory");
    }
    // small batches are latency-bound: the cooperative kernel's shorter per-env critical path wins
    // (cfg2 at 4,096 envs: 17.6 us vs 22.8 us); large batches of small domains are issue-bound and
    // the lane-per-env kernel executes a third of the instructions (cfg3 at 65,536: 20.7 vs 37 us)
    h->useCoop = h->coopFn && (!h->stepTile || cfg->B <= 8192);
    // register-resident kernel for the compile-time domains; it also emits the observations when
    // the observation tile leaves room for enough resident warps (cfg3: 22 KB per 32-env tile)
    // roles (warps) per 32-env tile.  Big batches of a small domain: every tile of a launch is resident
    // at once (12.4 KB of shared memory per tile), two roles give the schedulers twice the warps to
    // hide the serial latency of an environment (cfg3 at 65,536 envs, launches from independent shards
    // overlapping on 3 streams: 15.1 / 14.4 / 14.6 us per launch with 1 / 2 / 4 roles; strictly
    // serial launches: 21.9 / 23.8 / 24.6 us).  Small batches and bigger domains are bound by one
    // tile's critical path, which 4 roles cut (cfg2 domain at 65,536 envs: 24.6 us vs 41.8 us)
    {
        const size_t obsTile = (size_t)32 * lay.obs_halfs * 2;
        h->fusedRoles = (cfg->B >= 32768 && obsTile <= 28 * 1024) ? 2 : 4;
    }
    if (const char *e = getenv("MSCHED_ROLES")) { const int r = atoi(e); if (r == 1 || r == 2 || r == 4) h->fusedRoles = r; }
    h->fusedFn = pick_fused_kernel(cfg->N, cfg->C, cfg->L, h->fusedRoles);
    if (h->fusedFn && !getenv("MSCHED_NO_SPEC"))
        if (StepKernel sp = pick_fused_spec(*cfg, h->fusedRoles)) h->fusedFn = sp;
    if (h->fusedFn) {
        if (compute_compact_layout(cfg, lay, &h->clay) != MSCHED_OK) memset(&h->clay, 0, sizeof(h->clay));
        h->fusedSmem = fused_smem_bytes(lay.state_words, lay.action_halfs, lay.result_words, 0, cfg->C, h->clay.words);
        h->fusedSmemObs = fused_smem_bytes(lay.state_words, lay.action_halfs, lay.result_words, lay.obs_halfs, cfg->C, h->clay.words);
        h->fusedSmemNC = fused_smem_bytes(lay.state_words, lay.action_halfs, lay.result_words, 0, cfg->C, 0);
        h->fusedSmemObsNC = fused_smem_bytes(lay.state_words, lay.action_halfs, lay.result_words, lay.obs_halfs, cfg->C, 0);
        h->fuseObs = h->fusedSmemObs <= 48 * 1024;  // cfg2 domain: 43.6 KB tile, still one launch fewer
        if (const char *e = getenv("MSCHED_FUSE_OBS"))
            h->fuseObs = atoi(e) != 0 && h->fusedSmemObs + 2048 <= (size_t)h->smemOptin;
        if (h->fusedSmem + 64 > (size_t)h->smemOptin) h->fusedFn = nullptr;
    }
    if (h->fusedFn) {
        {   // resident CTAs per SM of a host-buffer (zero-copy) launch, see launch_step: MSCHED_HOST_CTAS, 0 = no limit
            int ctas = 3, smemSm = 0;
            if (const char *e = getenv("MSCHED_HOST_CTAS")) ctas = atoi(e);
            CUDA_TRY(cudaDeviceGetAttribute(&smemSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
            h->hostSmem = 0;
            if (ctas >= 1 && ctas <= 16) {
                long long v = (long long)smemSm / ctas - 1024 - 256;  // the per-CTA reserve and the static barriers
                if (v > h->smemOptin - 1024) v = h->smemOptin - 1024;
                h->hostSmem = v > 0 ? (size_t)(v & ~127ll) : 0;
            }
        }
        size_t fusedMax = h->fuseObs ? h->fusedSmemObs : h->fusedSmem;
        if (h->hostSmem > fusedMax) fusedMax = h->hostSmem;
        CUDA_TRY(cudaFuncSetAttribute(h->fusedFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fusedMax));
        h->useFused = true;
        h->useWarp = false;
        h->multiFn = h->fuseObs ? pick_fused_multi(*cfg, h->fusedRoles, !getenv("MSCHED_NO_SPEC")) : nullptr;
        if (h->multiFn)
            CUDA_TRY(cudaFuncSetAttribute(h->multiFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fusedSmemObs));
        h->hcFn = (h->fuseObs && !cfg->freePrices) ? pick_fused_hc(*cfg, h->fusedRoles, !getenv("MSCHED_NO_SPEC")) : nullptr;
        if (h->hcFn)
            CUDA_TRY(cudaFuncSetAttribute(h->hcFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fusedSmemObs));
    }
    // MSCHED_STEP_IMPL=fused|lane|coop forces one implementation (tests run all of them)
    if (const char *e = getenv("MSCHED_STEP_IMPL")) {
        const bool wantFused = !strcmp(e, "fused"), wantCoop = !strcmp(e, "coop"), wantLane = !strcmp(e, "lane"),
                   wantWarp = !strcmp(e, "warp");
        if ((wantFused && !h->fusedFn) || (wantCoop && !h->coopFn) || (wantLane && !h->stepTile) || (wantWarp && !h->warpSmem)) {
            return fail(MSCHED_E_ARG, "MSCHED_STEP_IMPL: that kernel is not available for this domain");
        }
        if (wantFused || wantCoop || wantLane || wantWarp) {
            h->useFused = wantFused;
            h->useCoop = wantCoop;
            h->useWarp = wantWarp;
        }
    }
    h->obsFn = pick_obs_kernel(cfg->N, cfg->C, cfg->L);
    if (h->obsFn && (lay.state_words * 4 > lay.obs_halfs * 2 || 32 * lay.obs_halfs * 2 + 64 > h->smemOptin))
        h->obsFn = nullptr;
    if (h->obsFn)
        CUDA_TRY(cudaFuncSetAttribute(h->obsFn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      32 * lay.obs_halfs * 2));
    CUDA_TRY(cudaMalloc(&h->stageAction, (size_t)lay.padded_envs * lay.action_halfs * 2));
    CUDA_TRY(cudaMalloc(&h->stageResult, (size_t)lay.padded_envs * lay.result_words * 4));
    CUDA_TRY(cudaMemset(h->stageAction, 0, (size_t)lay.padded_envs * lay.action_halfs * 2));
    CUDA_TRY(cudaMalloc(&h->roundDev, 2 * sizeof(int)));
    CUDA_TRY(cudaMemset(h->roundDev, 0, 2 * sizeof(int)));
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->hostStream[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->evDone[k], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&h->evStart, cudaEventDisableTiming));
    *handle = guard.release();
    return MSCHED_OK;
}

int msched_destroy(void *handle)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return MSCHED_OK;
    cudaSetDevice(h->device);
    free_handle(h);
    return MSCHED_OK;
}

int msched_get_info(void *handle, MschedInfo *out)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !out) return fail(MSCHED_E_ARG, "null handle/out");
    memset(out, 0, sizeof(*out));
    if (h->useFused) {
        out->step_impl = 2;
        out->fuses_observations = h->fuseObs ? 1 : 0;
        out->envs_per_cta = 32;
        out->threads_per_cta = 32 * h->fusedRoles;
        out->smem_bytes_per_cta = (int)(h->fuseObs ? h->fusedSmemObsNC : h->fusedSmemNC);
    } else if (h->useWarp) {
        out->step_impl = 3;
        out->envs_per_cta = 4;
        out->threads_per_cta = 128;
        out->smem_bytes_per_cta = (int)h->warpSmem;
    } else if (h->useCoop) {
        out->step_impl = 1;
        out->envs_per_cta = h->coopThreads / h->coopG;
        out->threads_per_cta = h->coopThreads;
        out->smem_bytes_per_cta = (int)h->coopSmem;
    } else {
        out->step_impl = 0;
        out->envs_per_cta = out->threads_per_cta = h->stepTile;
        out->smem_bytes_per_cta = (int)((size_t)h->stepTile * ((size_t)h->p.W * 4 + (size_t)h->p.AH * 2 +
                                                                (size_t)h->p.RW * 4 + (size_t)scratch_words(h->p.C) * 4));
    }
    return MSCHED_OK;
}

int msched_debug_timeline(void *handle, uint64_t *timeline_dev)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return fail(MSCHED_E_ARG, "null handle");
#ifdef MSCHED_TIMELINE
    h->p.timeline = reinterpret_cast<unsigned long long *>(timeline_dev);
    return MSCHED_OK;
#else
    if (timeline_dev) return fail(MSCHED_E_ARG, "the phase stamps are compiled out: rebuild with make NVEXTRA=-DMSCHED_TIMELINE");
    return MSCHED_OK;
#endif
}

int msched_bind_state(void *handle, void *state_dev, void *chain_dev)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !state_dev || !chain_dev) return fail(MSCHED_E_ARG, "null handle/state/chain");
    if (!aligned16(state_dev) || !aligned16(chain_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    h->p.state = static_cast<uint32_t *>(state_dev);
    h->p.chain = static_cast<uint32_t *>(chain_dev);
    return MSCHED_OK;
}

int msched_bind_stats(void *handle, int32_t *stats_dev)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return fail(MSCHED_E_ARG, "null handle");
    h->p.stats = stats_dev;  // NULL switches the statistics off
    return MSCHED_OK;
}

int msched_result_sums(void *handle, const uint32_t *result_dev, double *out_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !result_dev || !out_dev) return fail(MSCHED_E_ARG, "null handle/result/out");
    const int nOut = h->lay.result_words + 6;
    int blocks = (h->cfg.B + 255) / 256;
    if (blocks > 592) blocks = 592;
    result_sums_kernel<<<blocks, 256, nOut * sizeof(double), static_cast<cudaStream_t>(stream)>>>(h->p, result_dev, out_dev);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_stats_sums(void *handle, int64_t *out_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !out_dev) return fail(MSCHED_E_ARG, "null handle/out");
    if (!h->p.stats) return fail(MSCHED_E_STATE, "statistics buffer not bound");
    const int cols = 4 * h->cfg.J;
    dim3 block(32, 4), grid((h->cfg.B + 32 * 64 - 1) / (32 * 64), (cols + 3) / 4);
    if (grid.x < 1) grid.x = 1;
    stats_sums_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(h->p.stats, h->cfg.B, cols,
                                                                           reinterpret_cast<unsigned long long *>(out_dev));
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_reset(void *handle, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return fail(MSCHED_E_ARG, "null handle");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    CUDA_TRY(cudaSetDevice(h->device));
    h->round = 0;
    h->p.round = 0;
    CUDA_TRY(cudaMemsetAsync(h->roundDev, 0, sizeof(int), static_cast<cudaStream_t>(stream)));
    reset_kernel<<<(h->p.Bpad + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->p);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_get_round(void *handle, int64_t *round)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !round) return fail(MSCHED_E_ARG, "null handle/round");
    if (h->deviceRound) {  // graph replays advance the device counter only
        int r = 0;
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(&r, h->roundDev, sizeof(int), cudaMemcpyDeviceToHost));
        h->round = r;
    }
    *round = h->round;
    return MSCHED_OK;
}

int msched_set_round_mode(void *handle, int device_side, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return fail(MSCHED_E_ARG, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (device_side) {
        const int r = (int)h->round;
        CUDA_TRY(cudaMemcpyAsync(h->roundDev, &r, sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        h->deviceRound = true;
    } else if (h->deviceRound) {
        int64_t r = 0;
        int rc = msched_get_round(handle, &r);
        if (rc) return rc;
        h->deviceRound = false;
    }
    return MSCHED_OK;
}

int msched_set_round(void *handle, int64_t round)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || round < 0 || round > 0x7fffffff) return fail(MSCHED_E_ARG, "bad handle/round");
    h->round = round;
    if (h->deviceRound) {
        const int r = (int)round;
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaMemcpy(h->roundDev, &r, sizeof(int), cudaMemcpyHostToDevice));
    }
    return MSCHED_OK;
}

int msched_step(void *handle, const int16_t *action_dev, const double *spawn_u_dev, uint32_t *result_dev,
                void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !action_dev || !result_dev) return fail(MSCHED_E_ARG, "null handle/action/result");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(action_dev) || !aligned16(result_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64 && !spawn_u_dev) return fail(MSCHED_E_ARG, "spawn_u required");
    DevParams p = h->p;
    p.action = action_dev;
    p.spawnU = spawn_u_dev;
    p.result = result_dev;
    p.obs = nullptr;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    launch_step(h, p, static_cast<cudaStream_t>(stream));
    CUDA_TRY(cudaGetLastError());
    h->round += 1;
    return MSCHED_OK;
}

int msched_step_observe(void *handle, const int16_t *action_dev, const double *spawn_u_dev, uint32_t *result_dev,
                        int16_t *obs_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !obs_dev) return fail(MSCHED_E_ARG, "null handle/obs");
    if (!aligned16(obs_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    if (!(h->useFused && h->fuseObs)) {
        int rc = msched_step(handle, action_dev, spawn_u_dev, result_dev, stream);
        if (rc) return rc;
        return msched_observe_dense(handle, obs_dev, nullptr, stream);
    }
    if (!action_dev || !result_dev) return fail(MSCHED_E_ARG, "null action/result");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(action_dev) || !aligned16(result_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64 && !spawn_u_dev) return fail(MSCHED_E_ARG, "spawn_u required");
    DevParams p = h->p;
    p.action = action_dev;
    p.spawnU = spawn_u_dev;
    p.result = result_dev;
    p.obs = obs_dev;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    launch_step(h, p, static_cast<cudaStream_t>(stream));
    CUDA_TRY(cudaGetLastError());
    h->round += 1;
    return MSCHED_OK;
}

int msched_step_multi(void *handle, const int16_t *action_dev, int n_steps, uint32_t *result_dev, int16_t *obs_dev, int obs_every,
                      void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !action_dev || !result_dev) return fail(MSCHED_E_ARG, "null handle/action/result");
    if (n_steps < 1 || n_steps > 4096) return fail(MSCHED_E_ARG, "n_steps out of range");
    if (!h->useFused || !h->multiFn) return fail(MSCHED_E_ARG, "msched_step_multi: no multi-step kernel for this domain / role count (use msched_step_observe per step)");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64) return fail(MSCHED_E_ARG, "msched_step_multi: recorded float64 spawn draws are per step (use msched_step)");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(action_dev) || !aligned16(result_dev) || (obs_dev && !aligned16(obs_dev))) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    if (obs_every && !obs_dev) return fail(MSCHED_E_ARG, "obs_every needs an observation buffer");
    DevParams p = h->p;
    p.action = action_dev; p.spawnU = nullptr; p.result = result_dev; p.obs = obs_dev; p.cres = nullptr;
    p.nSteps = n_steps; p.obsEvery = obs_every ? 1 : 0;
    p.actStep = (long long)p.Bpad * p.AH; p.resStep = (long long)p.Bpad * p.RW; p.obsStep = (long long)p.Bpad * p.OH;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    p.roundDev = h->deviceRound ? h->roundDev : nullptr;
    p.roundTicket = h->deviceRound ? reinterpret_cast<unsigned *>(h->roundDev + 1) : nullptr;
    h->multiFn<<<p.Bpad / 32, 32 * h->fusedRoles, obs_dev ? h->fusedSmemObsNC : h->fusedSmemNC, static_cast<cudaStream_t>(stream)>>>(p);
    CUDA_TRY(cudaGetLastError());
    h->round += n_steps;
    return MSCHED_OK;
}

int msched_rollout_hardcoded(void *handle, int16_t *action_dev, int n_steps, uint32_t *result_dev, int16_t *obs_dev, int obs_every,
                             int random_ties, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !action_dev || !result_dev || !obs_dev) return fail(MSCHED_E_ARG, "null handle/action/result/obs");
    if (n_steps < 1 || n_steps > 4096) return fail(MSCHED_E_ARG, "n_steps out of range");
    if (!h->useFused || !h->hcFn) return fail(MSCHED_E_ARG, "msched_rollout_hardcoded: no multi-step kernel for this domain / role count");
    if (h->cfg.freePrices || h->cfg.auctionMode == MSCHED_AUCTION_EXTERNAL || h->cfg.spawnMode != MSCHED_SPAWN_PHILOX)
        return fail(MSCHED_E_ARG, "msched_rollout_hardcoded: fixed prices, in-kernel auction and device spawn draws only (the agents fill "
                                  "the acceptor and offer-core fields of the action record, nothing else)");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(action_dev) || !aligned16(result_dev) || !aligned16(obs_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    DevParams p = h->p;
    p.action = action_dev; p.actionOut = action_dev; p.spawnU = nullptr; p.result = result_dev; p.obs = obs_dev; p.cres = nullptr;
    p.nSteps = n_steps; p.obsEvery = obs_every ? 1 : 0;
    p.actStep = 0; p.resStep = (long long)p.Bpad * p.RW; p.obsStep = (long long)p.Bpad * p.OH;
    p.hcPolicy = 1; p.hcRandomTies = random_ties ? 1 : 0;
    p.hcOAcc = h->lay.o_acceptor; p.hcOOff = h->lay.o_offer; p.hcAccRow = h->lay.o_acc_row; p.hcOffRow = h->lay.o_off_row;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    p.roundDev = h->deviceRound ? h->roundDev : nullptr;
    p.roundTicket = h->deviceRound ? reinterpret_cast<unsigned *>(h->roundDev + 1) : nullptr;
    h->hcFn<<<p.Bpad / 32, 32 * h->fusedRoles, h->fusedSmemObsNC, static_cast<cudaStream_t>(stream)>>>(p);
    CUDA_TRY(cudaGetLastError());
    h->round += n_steps;
    return MSCHED_OK;
}

int msched_get_compact_result_layout(const MschedConfig *cfg, MschedCompactResultLayout *out)
{
    MschedLayout lay;
    int rc = compute_layout(cfg, &lay);
    if (rc) return rc;
    return compute_compact_layout(cfg, lay, out);
}

int msched_step_host_compact(void *handle, const int16_t *action_host, uint32_t *cresult_host, int16_t *obs_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !action_host || !cresult_host) return fail(MSCHED_E_ARG, "null handle/action/result");
    if (!h->useFused || !h->clay.words) return fail(MSCHED_E_ARG, "compact results need a fused-kernel domain with half-exact rewards (use msched_step_host)");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64) return fail(MSCHED_E_ARG, "step_host does not take recorded draws");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (h->cfg.B != h->lay.padded_envs || !aligned16(action_host) || !aligned16(cresult_host) || (obs_dev && !aligned16(obs_dev)))
        return fail(MSCHED_E_ARG, "compact host step: B must be a multiple of 128 and the buffers 16-byte aligned");
    cudaPointerAttributes aa{}, ra{};
    if (cudaPointerGetAttributes(&aa, action_host) != cudaSuccess || aa.type != cudaMemoryTypeHost || !aa.devicePointer ||
        cudaPointerGetAttributes(&ra, cresult_host) != cudaSuccess || ra.type != cudaMemoryTypeHost || !ra.devicePointer) {
        (void)cudaGetLastError();
        return fail(MSCHED_E_ARG, "compact host step: the host buffers must be pinned (cudaHostAlloc / pin_memory)");
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(h->device));
    const bool fuse = obs_dev && h->fuseObs;
    DevParams p = h->p;
    p.action = static_cast<const int16_t *>(aa.devicePointer);
    p.result = nullptr;
    p.cres = static_cast<uint32_t *>(ra.devicePointer);
    p.CW = h->clay.words;
    p.cOffer = h->clay.c_offer; p.cPrice = h->clay.c_price; p.cAcc = h->clay.c_acceptor; p.cAuc = h->clay.c_auctioneer;
    p.cAgent = h->clay.c_agent; p.cTail = h->clay.c_quality;
    p.spawnU = nullptr;
    p.obs = fuse ? obs_dev : nullptr;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    launch_step(h, p, s, true, true);
    CUDA_TRY(cudaGetLastError());
    h->round += 1;
    if (obs_dev && !fuse) {
        int rc = msched_observe_dense(handle, obs_dev, nullptr, stream);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return MSCHED_OK;
}

int msched_observe_compact(void *handle, int16_t *cobs_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !cobs_dev) return fail(MSCHED_E_ARG, "null handle/cobs");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    DevParams p = h->p;
    p.cobs = cobs_dev;
    launch_observe_compact(p, static_cast<cudaStream_t>(stream));
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_step_compact(void *handle, const int16_t *action_dev, const double *spawn_u_dev, uint32_t *result_dev,
                        int16_t *cobs_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !cobs_dev) return fail(MSCHED_E_ARG, "null handle/cobs");
    if (!h->useWarp) {
        int rc = msched_step(handle, action_dev, spawn_u_dev, result_dev, stream);
        if (rc) return rc;
        return msched_observe_compact(handle, cobs_dev, stream);
    }
    if (!action_dev || !result_dev) return fail(MSCHED_E_ARG, "null action/result");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(action_dev) || !aligned16(result_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64 && !spawn_u_dev) return fail(MSCHED_E_ARG, "spawn_u required");
    DevParams p = h->p;
    p.action = action_dev;
    p.spawnU = spawn_u_dev;
    p.result = result_dev;
    p.obs = nullptr;
    p.cobs = cobs_dev;
    p.round = (int)h->round;
    p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
    launch_step(h, p, static_cast<cudaStream_t>(stream));
    CUDA_TRY(cudaGetLastError());
    h->round += 1;
    return MSCHED_OK;
}

int msched_step_host(void *handle, const int16_t *action_host, uint32_t *result_host, int16_t *obs_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !action_host || !result_host) return fail(MSCHED_E_ARG, "null handle/action/result");
    if (h->cfg.spawnMode == MSCHED_SPAWN_U64) return fail(MSCHED_E_ARG, "step_host does not take recorded draws");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (obs_dev && !aligned16(obs_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(h->device));
    const bool fuse = obs_dev && h->useFused && h->fuseObs;
    // Zero-copy path: when both host buffers are pinned (device-accessible under unified addressing) and the
    // batch has no padding, the fused kernel's TMA bulk copies read the action tiles from and write the result
    // tiles to HOST memory directly -- one launch, no staging copies, and the PCIe traffic of the tiles
    // pipelines across the resident CTAs (reads of late tiles overlap writes of early ones) instead of
    // running as three serial phases per chunk.  MSCHED_HOST_ZEROCOPY=0 keeps the staged path.
    if (h->useFused && h->cfg.B == h->lay.padded_envs && aligned16(action_host) && aligned16(result_host)) {
        // MSCHED_HOST_ZEROCOPY: 1 both directions, 2 results only (actions staged by the copy engine), 3 actions only
        // (results staged and copied out by the copy engine, chunked so that copy k overlaps kernel k+1), 0 off
        static const int mode = [] { const char *e = getenv("MSCHED_HOST_ZEROCOPY"); return e ? atoi(e) : 1; }();
        cudaPointerAttributes aa{}, ra{};
        if (mode > 0 && cudaPointerGetAttributes(&aa, action_host) == cudaSuccess && aa.type == cudaMemoryTypeHost && aa.devicePointer &&
            cudaPointerGetAttributes(&ra, result_host) == cudaSuccess && ra.type == cudaMemoryTypeHost && ra.devicePointer) {
            const int B = h->cfg.B, AH = h->lay.action_halfs, RW = h->lay.result_words;
            int nChunks = 1;
            if (const char *e = getenv("MSCHED_HOST_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= 64) nChunks = v; }
            const int chunk = ((B + nChunks - 1) / nChunks + MSCHED_TILE_ENVS - 1) / MSCHED_TILE_ENVS * MSCHED_TILE_ENVS;
            const bool multi = chunk < B || mode >= 2;
            if (multi) {
                CUDA_TRY(cudaEventRecord(h->evStart, s));
                for (int k = 0; k < 2; ++k) CUDA_TRY(cudaStreamWaitEvent(h->hostStream[k], h->evStart, 0));
            }
            int c = 0;
            for (int e0 = 0; e0 < B; e0 += chunk, ++c) {
                const int n = (B - e0 < chunk) ? (B - e0) : chunk;
                cudaStream_t cs = multi ? h->hostStream[c & 1] : s;
                DevParams p = h->p;
                p.B = n;
                p.Bpad = n;  // n is a multiple of the padding unit here
                p.state = h->p.state + (size_t)e0 * p.W;
                p.chain = h->p.chain + (size_t)e0 * h->lay.chain_words;
                p.stats = h->p.stats ? h->p.stats + (size_t)e0 * h->cfg.J * 4 : nullptr;
                if (mode == 2) {
                    CUDA_TRY(cudaMemcpyAsync(h->stageAction + (size_t)e0 * AH, action_host + (size_t)e0 * AH, (size_t)n * AH * 2,
                                             cudaMemcpyHostToDevice, cs));
                    p.action = h->stageAction + (size_t)e0 * AH;
                } else {
                    p.action = static_cast<const int16_t *>(aa.devicePointer) + (size_t)e0 * AH;
                }
                p.result = mode == 3 ? h->stageResult + (size_t)e0 * RW : static_cast<uint32_t *>(ra.devicePointer) + (size_t)e0 * RW;
                p.spawnU = nullptr;
                p.obs = fuse ? obs_dev + (size_t)e0 * h->lay.obs_halfs : nullptr;
                p.envOffset = h->p.envOffset + e0;
                p.round = (int)h->round;
                p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
                launch_step(h, p, cs, !multi, true);
                CUDA_TRY(cudaGetLastError());
                if (mode == 3)
                    CUDA_TRY(cudaMemcpyAsync(result_host + (size_t)e0 * RW, h->stageResult + (size_t)e0 * RW, (size_t)n * RW * 4,
                                             cudaMemcpyDeviceToHost, cs));
            }
            if (multi) {
                for (int k = 0; k < 2; ++k) {
                    CUDA_TRY(cudaEventRecord(h->evDone[k], h->hostStream[k]));
                    CUDA_TRY(cudaStreamWaitEvent(s, h->evDone[k], 0));
                }
                if (h->deviceRound) bump_round_kernel<<<1, 1, 0, s>>>(h->roundDev);
            }
            h->round += 1;
            if (obs_dev && !fuse) {
                int rc = msched_observe_dense(handle, obs_dev, nullptr, stream);
                if (rc) return rc;
            }
            CUDA_TRY(cudaStreamSynchronize(s));
            return MSCHED_OK;
        }
        (void)cudaGetLastError();  // pageable memory: not an error, take the staged path
    }
    // The batch is cut into chunks (multiples of the 128-env padding unit) that alternate between
    // two internal streams, so the H2D copy of one chunk, the kernel of another and the D2H copy of
    // a third overlap (PCIe is full duplex); environments are independent, so any split is exact.
    const int B = h->cfg.B, AH = h->lay.action_halfs, RW = h->lay.result_words;
    int nChunks = 2;
    if (const char *e = getenv("MSCHED_HOST_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= 64) nChunks = v; }
    int chunk = ((B + nChunks - 1) / nChunks + MSCHED_TILE_ENVS - 1) / MSCHED_TILE_ENVS * MSCHED_TILE_ENVS;
    CUDA_TRY(cudaEventRecord(h->evStart, s));
    for (int k = 0; k < 2; ++k) CUDA_TRY(cudaStreamWaitEvent(h->hostStream[k], h->evStart, 0));
    int c = 0;
    for (int e0 = 0; e0 < B; e0 += chunk, ++c) {
        const int n = (B - e0 < chunk) ? (B - e0) : chunk;
        cudaStream_t cs = h->hostStream[c & 1];
        CUDA_TRY(cudaMemcpyAsync(h->stageAction + (size_t)e0 * AH, action_host + (size_t)e0 * AH, (size_t)n * AH * 2,
                                 cudaMemcpyHostToDevice, cs));
        DevParams p = h->p;
        p.B = n;
        p.Bpad = msched_padded_envs(n);
        p.state = h->p.state + (size_t)e0 * p.W;
        p.chain = h->p.chain + (size_t)e0 * h->lay.chain_words;
        p.stats = h->p.stats ? h->p.stats + (size_t)e0 * h->cfg.J * 4 : nullptr;
        p.action = h->stageAction + (size_t)e0 * AH;
        p.result = h->stageResult + (size_t)e0 * RW;
        p.spawnU = nullptr;
        p.obs = fuse ? obs_dev + (size_t)e0 * h->lay.obs_halfs : nullptr;
        p.envOffset = h->p.envOffset + e0;
        p.round = (int)h->round;
        p.doneFlag = ((h->round + 1) % h->cfg.episodeLength) == 0 ? 1 : 0;
        launch_step(h, p, cs, false);  // several launches per step: the round advances once, below
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(result_host + (size_t)e0 * RW, h->stageResult + (size_t)e0 * RW, (size_t)n * RW * 4,
                                 cudaMemcpyDeviceToHost, cs));
    }
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(cudaEventRecord(h->evDone[k], h->hostStream[k]));
        CUDA_TRY(cudaStreamWaitEvent(s, h->evDone[k], 0));
    }
    if (h->deviceRound) bump_round_kernel<<<1, 1, 0, s>>>(h->roundDev);
    h->round += 1;
    if (obs_dev && !fuse) {  // two-launch domains: the observation kernel follows on the caller's stream
        int rc = msched_observe_dense(handle, obs_dev, nullptr, stream);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return MSCHED_OK;
}

int msched_observe_dense(void *handle, int16_t *obs_dev, int16_t *ids_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !obs_dev) return fail(MSCHED_E_ARG, "null handle/obs");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (!aligned16(obs_dev)) return fail(MSCHED_E_ARG, "buffers must be 16-byte aligned");
    DevParams p = h->p;
    p.obs = obs_dev;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (h->obsFn)
        h->obsFn<<<p.Bpad / 32, 32, (size_t)32 * p.OH * 2, s>>>(p);
    else
        observe_kernel_direct<<<(p.Bpad + 63) / 64, 64, 0, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    if (ids_dev) {
        ids_kernel<<<(p.B + 63) / 64, 64, 0, s>>>(p, ids_dev);
        CUDA_TRY(cudaGetLastError());
    }
    return MSCHED_OK;
}

int msched_auctioneer_action(void *handle, int random_ties, int16_t *out_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !out_dev) return fail(MSCHED_E_ARG, "null handle/out");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    DevParams p = h->p;
    p.round = (int)h->round;
    p.roundDev = h->deviceRound ? h->roundDev : nullptr;
    auctioneer_kernel<<<(p.B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p, random_ties, out_dev);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_hardcoded_actions(void *handle, const int16_t *obs_dev, int random_ties, const float *u_override_dev,
                             int16_t *action_dev, int32_t *ncand_dev, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h || !obs_dev || !action_dev) return fail(MSCHED_E_ARG, "null handle/obs/action");
    DevParams p = h->p;
    p.round = (int)h->round;
    p.roundDev = h->deviceRound ? h->roundDev : nullptr;
    HardcodedArgs a;
    a.obs = obs_dev; a.action = action_dev; a.uOverride = u_override_dev; a.ncand = ncand_dev;
    a.oAcc = h->lay.o_acceptor; a.oOff = h->lay.o_offer; a.accRow = h->lay.o_acc_row; a.offRow = h->lay.o_off_row;
    a.randomTies = random_ties;
    const long long M = (long long)p.B * (p.N * p.C + p.NL);
    hardcoded_policy_kernel<<<(unsigned)((M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p, a);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_export_state(void *handle, int env0, int count, int32_t *core, int32_t *slot, int32_t *offer,
                        int32_t *chain, int32_t *chain_len, int32_t *misc, void *stream)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h) return fail(MSCHED_E_ARG, "null handle");
    if (!h->p.state) return fail(MSCHED_E_STATE, "state not bound");
    if (env0 < 0 || count < 0 || env0 + count > h->cfg.B) return fail(MSCHED_E_ARG, "env range out of bounds");
    if (count == 0) return MSCHED_OK;
    ExportArgs a{env0, count, core, slot, offer, chain, chain_len, misc};
    export_kernel<<<(count + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(h->p, a);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

}  // extern "C"
