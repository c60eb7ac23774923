// policy_kernels.cuh -- the PPO pieces of the rollout path.
//
//   src/PPOmodules.py:32-39,53-63   ActorCritic.actor + act(): Linear-Tanh-Linear-Tanh-Linear-
//                                   Softmax, Categorical.sample, Categorical.log_prob
//   src/PPOmodules.py:114-125       PPO.selectAction (state.float())
//   src/PPOmodules.py:128-137       PPO.update returns prologue
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include "msched_common.cuh"
#include "policy_common.cuh"

namespace msched {

#ifdef MSCHED_ACTOR_DISPATCH_TU  // (plain kernels: defined once, in the unit that launches them)
// ---- discounted returns --------------------------------------------------------------------
// One lane per unit m; rewards/out are [T][M] so that a warp touches 128 contiguous bytes per
// time step.  G_t = r_t + gamma*G_{t+1} in float64 like the Python loop, cast to float32; the
// normalisation statistics are accumulated in float64 over the float32 values in the same sweep
// and a second sweep recomputes G instead of re-reading it (12 B/element of HBM traffic).
// x / d for one divisor d and many x: the quotient from the reciprocal (computed once per column) plus one residual
// correction -- q = x * (1/d); q += (x - q d) * (1/d) -- which is the correctly rounded quotient except in rare last-bit
// ties (the inline IEEE division is ~14 instructions per element and was 56 % of the tensor-map kernel's instructions)
__device__ __forceinline__ float div_by(float x, float d, float inv)
{
    const float q = x * inv;
    return fmaf(fmaf(-q, d, x), inv, q);
}

__global__ void returns_kernel(const float *__restrict__ r, int T, int M, double gamma, int normalise,
                               float *__restrict__ out)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double disc = 0.0, s1 = 0.0, s2 = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        disc = __dadd_rn((double)r[(size_t)t * M + m], __dmul_rn(gamma, disc));
        const float g = (float)disc;
        if (!normalise) out[(size_t)t * M + m] = g;
        s1 += (double)g;
        s2 += (double)g * (double)g;
    }
    if (!normalise) return;
    const float mean = (float)(s1 / T);
    // unbiased variance about the float32 mean, from the float64 moments
    const double dm = (double)mean;
    double var = (s2 - 2.0 * dm * s1 + (double)T * dm * dm) / (double)(T - 1);
    if (var < 0.0) var = 0.0;
    const float sd = (float)sqrt(var);
    const float denom = sd + 1e-7f, inv = 1.0f / denom;
    disc = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        disc = __dadd_rn((double)r[(size_t)t * M + m], __dmul_rn(gamma, disc));
        out[(size_t)t * M + m] = div_by((float)disc - mean, denom, inv);
    }
}

// ---- discounted returns, TMA-tiled -----------------------------------------------------------------
// The same computation with every reward read from HBM ONCE: a CTA owns W consecutive units; one thread issues a
// 4W-byte bulk copy (cp.async.bulk, 1-D TMA) per time step, all T of them in flight at once on one mbarrier, so
// the whole [T][128] reward tile (100 KB at T = 200) lands in shared memory at full memory-level parallelism.
// Each thread then scans its column backwards in float64 like the Python loop, keeps G in place of r, forms the
// moments, normalises in place, and the tile leaves with T bulk stores.  8 B/element of HBM traffic instead of 12
// and no second float64 chain.  Needs M % 4 == 0 (16-byte rows) and T * 512 B of shared memory.
template <int W>
__global__ void __launch_bounds__(W) returns_tile_kernel(const float *__restrict__ r, int T, int M, double gamma, int normalise,
                                                           float *__restrict__ out)
{
    extern __shared__ __align__(128) float tile[];  // [T][W]
    __shared__ __align__(8) uint64_t bar;
    const int m0 = blockIdx.x * W, tid = threadIdx.x;
    const int cols = min(W, M - m0);  // a multiple of 4
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, (uint32_t)T * (uint32_t)cols * 4u);
    }
    __syncthreads();
    // every thread issues its share of the T row copies (one thread issuing all of them is a serial chain of T issues)
    for (int t = tid; t < T; t += W) bulk_g2s(tile + t * W, r + (size_t)t * M + m0, (uint32_t)cols * 4u, &bar);
    mbar_wait(&bar, 0);
    if (tid < cols) {
        double disc = 0.0, s1 = 0.0, s2 = 0.0;
        for (int t = T - 1; t >= 0; --t) {
            disc = __dadd_rn((double)tile[t * W + tid], __dmul_rn(gamma, disc));
            const float g = (float)disc;
            tile[t * W + tid] = g;
            s1 += (double)g;
            s2 += (double)g * (double)g;
        }
        if (normalise) {
            const float mean = (float)(s1 / T);
            const double dm = (double)mean;
            double var = (s2 - 2.0 * dm * s1 + (double)T * dm * dm) / (double)(T - 1);
            if (var < 0.0) var = 0.0;
            const float denom = (float)sqrt(var) + 1e-7f, inv = 1.0f / denom;
#pragma unroll 8
            for (int t = 0; t < T; ++t) tile[t * W + tid] = div_by(tile[t * W + tid] - mean, denom, inv);
        }
    }
    fence_async_smem();
    __syncthreads();
    for (int t = tid; t < T; t += W) bulk_s2g(out + (size_t)t * M + m0, tile + t * W, (uint32_t)cols * 4u);
    bulk_commit();
    bulk_wait_read();
}

// The tile moved by ONE 2-D tensor-map copy each way (cp.async.bulk.tensor.2d, box [T][W], T <= 256): the 1-D version
// above issues 2 T row copies of 512 bytes per tile, and the copy unit's fixed time per copy -- not the bytes -- is what
// bounds it (1.2 M copies per call at T = 200, 393,216 units).  Columns beyond M are filled with zeros on the way in and
// dropped on the way out by the copy unit.  The wait is bounded: a descriptor the hardware rejects traps instead of hanging
template <int W>
__global__ void __launch_bounds__(W) returns_tmap_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                                                           int T, int M, double gamma, int normalise)
{
    extern __shared__ __align__(128) float tile[];  // [T][W]
    __shared__ __align__(8) uint64_t bar;
    const int m0 = blockIdx.x * W, tid = threadIdx.x;
    const int cols = min(W, M - m0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_async_smem();
        mbar_expect_tx(&bar, (uint32_t)T * (uint32_t)W * 4u);  // the whole box counts, clipped or not
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(&tin), "r"(m0), "r"(0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    {
        uint32_t ok = 0u;
        for (int it = 0; it < (1 << 22) && !ok; ++it)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (!ok) __trap();
    }
    if (tid < cols) {
        double disc = 0.0, s1 = 0.0, s2 = 0.0;
        for (int t = T - 1; t >= 0; --t) {
            disc = __dadd_rn((double)tile[t * W + tid], __dmul_rn(gamma, disc));
            const float g = (float)disc;
            tile[t * W + tid] = g;
            s1 += (double)g;
            s2 += (double)g * (double)g;
        }
        if (normalise) {
            const float mean = (float)(s1 / T);
            const double dm = (double)mean;
            double var = (s2 - 2.0 * dm * s1 + (double)T * dm * dm) / (double)(T - 1);
            if (var < 0.0) var = 0.0;
            const float denom = (float)sqrt(var) + 1e-7f, inv = 1.0f / denom;
#pragma unroll 8
            for (int t = 0; t < T; ++t) tile[t * W + tid] = div_by(tile[t * W + tid] - mean, denom, inv);
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(&tout), "r"(m0), "r"(0), "r"(smem_u32(tile)) : "memory");
        bulk_commit();
        bulk_wait_read();
    }
}

#endif  // MSCHED_ACTOR_DISPATCH_TU

// ---- actor forward, fp32 SIMT version --------------------------------------------------------
// grid = (ceil(n_envs / 128), units); a CTA evaluates ONE unit (one net) for 128 consecutive
// environments, so every lane reads the same weight at the same time (shared-memory broadcast,
// 128-bit) and FFMA is the bound.  Weights are staged transposed ([in][out]) in shared memory.
// lg[] are BASE-2 logits (the callers fold log2(e) into the last layer's weights and bias when they
// stage them), so the softmax needs one ex2 per action; lg[o] must be -inf for o >= A (padded
// bias), so every sweep runs unpredicated over the AP registers
template <int AP>
__device__ __forceinline__ int actor_epilogue(const ActorArgs &a, float (&lg)[AP], int A, long long row, int env,
                                              int unit, int gsel)
{
    float mx = lg[0];
#pragma unroll
    for (int o = 1; o < AP; ++o) mx = fmaxf(mx, lg[o]);
    float sum = 0.f;
#pragma unroll
    for (int o = 0; o < AP; ++o) { lg[o] = ex2_approx(lg[o] - mx); sum += lg[o]; }
    const float inv = 1.f / sum;
    float tot = 0.f;  // Categorical(probs) renormalises by the sum of the softmax output
#pragma unroll
    for (int o = 0; o < AP; ++o) { lg[o] *= inv; tot += lg[o]; }
    if (a.probs) {
#pragma unroll
        for (int o = 0; o < AP; ++o)
            if (o < A) a.probs[(size_t)row * A + o] = lg[o];
    }
    if (!a.action && !a.logprob && !a.actionRec) return -1;
    float u;
    if (a.uOverride) {
        u = a.uOverride[row];
    } else {
        const unsigned long long g = (unsigned long long)(a.rowOffset + row);
        uint32_t x4[4];
        const unsigned long long stp = a.stepDev ? *a.stepDev : a.step;
        philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stp,
                      (kStreamPolicy << 28) | (uint32_t)((stp >> 32) & 0x0fffffffu), (uint32_t)a.seed,
                      (uint32_t)(a.seed >> 32), x4);
        u = (float)(x4[0] >> 8) * (1.0f / 16777216.0f);
    }
    // inverse CDF: the first action whose cumulative probability exceeds u * total; padded entries
    // add 0 and can never be the first to exceed
    const float thr = u * tot;
    float cdf = 0.f, pa = 0.f;
    int act = -1;
#pragma unroll
    for (int o = 0; o < AP; ++o) {
        cdf += lg[o];
        const bool hit = act < 0 && cdf > thr;
        act = hit ? o : act;
        pa = hit ? lg[o] : pa;
    }
    if (act < 0) {  // u * total rounded up to the total: the last action
        act = A - 1;
#pragma unroll
        for (int o = 0; o < AP; ++o) pa = (o == A - 1) ? lg[o] : pa;
    }
    if (a.action) a.action[row] = act;  // what PPO.selectAction stores in buffer.actions
    if (a.actionRec)                    // what the world is handed (price -5 next to core action 0)
        a.actionRec[(size_t)env * a.actionRecStride + unit] = (int16_t)(gsel == 0 ? -5 : act);  // gsel is -1 without a gather
    if (a.logprob) {
        const float eps = 1.1920928955078125e-07f;
        float pn = pa / tot;
        pn = fminf(fmaxf(pn, eps), 1.f - eps);
        a.logprob[row] = logf(pn);
    }
    return act;
}

// shared-memory image of one actor net for the SIMT kernels: W1t [nIn][H] | b1 [H] | W2t [H][H] |
// b2 [H] | W3t [H][Apad] | b3 [Apad]; weights transposed ([in][out]), the last layer scaled by log2(e)
// (base-2 logits for the epilogue), padded actions get bias -inf
template <int H>
struct SimtNet {
    float *W1t, *b1, *W2t, *b2, *W3t, *b3;
    int nIn, A, Apad;
    __device__ __forceinline__ static int floats(int nIn, int A) { const int Ap = (A + 3) & ~3; return nIn * H + H + H * H + H + H * Ap + Ap; }
    __device__ __forceinline__ void carve(float *base, int nIn_, int A_)
    {
        nIn = nIn_; A = A_; Apad = (A_ + 3) & ~3;
        W1t = base; b1 = W1t + nIn * H; W2t = b1 + H; b2 = W2t + H * H; W3t = b2 + H; b3 = W3t + H * Apad;
    }
    // called by all 128 threads; a warp reads one weight row (contiguous) at a time.  Needs a
    // __syncthreads() between zero_pad() and stage(), and one after stage()
    __device__ __forceinline__ void zero_pad() { for (int i = threadIdx.x; i < H * Apad; i += blockDim.x) W3t[i] = 0.f; }
    __device__ __forceinline__ void stage(const float *w)
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const float *w2 = w + H * nIn + H, *w3 = w2 + H * H + H;
        for (int o = warp; o < H; o += 4)
            for (int k = lane; k < nIn; k += 32) W1t[k * H + o] = w[o * nIn + k];
        for (int o = warp; o < H; o += 4)
            for (int k = lane; k < H; k += 32) W2t[k * H + o] = w2[o * H + k];
        for (int i = threadIdx.x; i < H; i += blockDim.x) { b1[i] = w[H * nIn + i]; b2[i] = w2[H * H + i]; }
        for (int o = warp; o < A; o += 4)
            for (int k = lane; k < H; k += 32) W3t[k * Apad + o] = w3[o * H + k] * kLog2e;
        for (int i = threadIdx.x; i < Apad; i += blockDim.x) b3[i] = i < A ? w3[A * H + i] * kLog2e : -INFINITY;
    }
    // one row: xf(k) is input k; the inputs are fetched in batches of 16 (one memory round trip
    // per 16 inputs instead of one per input)
    template <int AP, class XF>
    __device__ __forceinline__ void forward(XF xf, float (&lg)[AP]) const
    {
        float h1[H], h2[H];
#pragma unroll
        for (int o = 0; o < H; ++o) h1[o] = b1[o];
        for (int k0 = 0; k0 < nIn; k0 += 16) {
            float xb[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) xb[q] = k0 + q < nIn ? xf(k0 + q) : 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (k0 + q < nIn) {
                    const float xv = xb[q];
                    const float4 *wr = reinterpret_cast<const float4 *>(W1t + (k0 + q) * H);
#pragma unroll
                    for (int o4 = 0; o4 < H / 4; ++o4) {
                        const float4 wv = wr[o4];
                        h1[4 * o4 + 0] = fmaf(wv.x, xv, h1[4 * o4 + 0]);
                        h1[4 * o4 + 1] = fmaf(wv.y, xv, h1[4 * o4 + 1]);
                        h1[4 * o4 + 2] = fmaf(wv.z, xv, h1[4 * o4 + 2]);
                        h1[4 * o4 + 3] = fmaf(wv.w, xv, h1[4 * o4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < H; ++o) { h1[o] = fast_tanh(h1[o]); h2[o] = b2[o]; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const float xv = h1[k];
            const float4 *wr = reinterpret_cast<const float4 *>(W2t + k * H);
#pragma unroll
            for (int o4 = 0; o4 < H / 4; ++o4) {
                const float4 wv = wr[o4];
                h2[4 * o4 + 0] = fmaf(wv.x, xv, h2[4 * o4 + 0]);
                h2[4 * o4 + 1] = fmaf(wv.y, xv, h2[4 * o4 + 1]);
                h2[4 * o4 + 2] = fmaf(wv.z, xv, h2[4 * o4 + 2]);
                h2[4 * o4 + 3] = fmaf(wv.w, xv, h2[4 * o4 + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < H; ++o) h2[o] = fast_tanh(h2[o]);
#pragma unroll
        for (int o = 0; o < AP; ++o) lg[o] = -INFINITY;
#pragma unroll
        for (int o4 = 0; o4 < AP / 4; ++o4) {
            if (4 * o4 < A) {
                float4 acc = *reinterpret_cast<const float4 *>(b3 + 4 * o4);
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    const float4 wv = *reinterpret_cast<const float4 *>(W3t + k * Apad + 4 * o4);
                    acc.x = fmaf(wv.x, h2[k], acc.x);
                    acc.y = fmaf(wv.y, h2[k], acc.y);
                    acc.z = fmaf(wv.z, h2[k], acc.z);
                    acc.w = fmaf(wv.w, h2[k], acc.w);
                }
                lg[4 * o4 + 0] = acc.x; lg[4 * o4 + 1] = acc.y; lg[4 * o4 + 2] = acc.z; lg[4 * o4 + 3] = acc.w;
            }
        }
    }
};

template <int H, int AP>
__global__ void __launch_bounds__(128) actor_forward_simt(const ActorArgs a)
{
    extern __shared__ __align__(16) float sw[];
    const int nIn = a.nIn, A = a.nActions;
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + H * H + H + A * H + A;
    SimtNet<H> n1;
    n1.carve(sw, nIn, A);
    n1.zero_pad();
    __syncthreads();
    n1.stage(a.weights + (size_t)net * pc);
    __syncthreads();

    // persistent: this CTA's share of the unit's 128-environment tiles
    const int nTiles = (a.nEnvs + 127) / 128;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int env = tile * 128 + threadIdx.x;
        if (env >= a.nEnvs) continue;
        const int16_t *xr = a.x + (size_t)env * a.envStride + (size_t)unit * a.unitStride;
        const long long row = (long long)env * a.units + unit;
        // FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): the price chooser's input is
        // [core prio, core rem, slot prio, slot rem] of the core the core chooser picked, sliced out
        // of the offer observation row; core action 0 feeds the dummy [-5,-5,-5,-5] (quirk Q1)
        int gsel = -1;
        int16_t g4[4] = {0, 0, 0, 0};
        if (a.gatherCore) {
            gsel = a.gatherCore[row];
            const int c2 = 2 * a.nCores;
            if (gsel <= 0 || gsel > a.nCores) {
                g4[0] = g4[1] = g4[2] = g4[3] = (int16_t)-5;
            } else {
                g4[0] = xr[2 * gsel]; g4[1] = xr[2 * gsel + 1]; g4[2] = xr[c2]; g4[3] = xr[c2 + 1];
            }
        }
        if (a.xUsed)
            for (int k = 0; k < nIn; ++k) a.xUsed[(size_t)row * nIn + k] = a.gatherCore ? g4[k & 3] : xr[k];
        float lg[AP];
        if (a.gatherCore)
            n1.template forward<AP>([&](int k) { return (float)g4[k & 3]; }, lg);
        else
            n1.template forward<AP>([&](int k) { return (float)xr[k]; }, lg);
        actor_epilogue(a, lg, A, row, env, unit, gsel);
    }
}

#ifdef MSCHED_ACTOR_DISPATCH_TU  // (plain kernels: defined once, in the unit that launches them)
// ---- DQNEntity.selectAction (src/DQNmodules.py:34-76), batched ------------------------------------
// Q-net Linear(in,16)-Tanh-Linear(16,A); epsilon-greedy: with probability epsilon (the caller's
// RUN_END + (RUN_START-RUN_END)*exp(-round/RUN_DECAY)) a uniformly random action, else arg-max Q
// (first maximum, like torch.max).  Weights per net: [W1 16*in | b1 16 | W2 A*16 | b2 A], torch layout.
// One thread per (environment, unit) row, the unit's weights staged in shared memory.
struct QArgs {
    ActorArgs a;     // x / strides / units / nets / seed / step / uOverride ([M][2]) / action / actionRec
    float epsilon;
    float *qOut;     // [M][A] or null
};

__global__ void __launch_bounds__(128) dqn_select_kernel(const QArgs q)
{
    constexpr int H = 16;
    extern __shared__ __align__(16) float sq[];
    const ActorArgs &a = q.a;
    const int nIn = a.nIn, A = a.nActions;
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + A * H + A;
    const float *w = a.weights + (size_t)net * pc;
    float *W1t = sq, *b1 = W1t + nIn * H, *W2 = b1 + H, *b2 = W2 + A * H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < H; o += 4)
        for (int k = lane; k < nIn; k += 32) W1t[k * H + o] = w[o * nIn + k];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = w[H * nIn + i];
    const float *w2 = w + H * nIn + H;
    for (int i = threadIdx.x; i < A * H; i += blockDim.x) W2[i] = w2[i];
    for (int i = threadIdx.x; i < A; i += blockDim.x) b2[i] = w2[A * H + i];
    __syncthreads();
    const int nTiles = (a.nEnvs + 127) / 128;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int env = tile * 128 + threadIdx.x;
        if (env >= a.nEnvs) continue;
        const int16_t *xr = a.x + (size_t)env * a.envStride + (size_t)unit * a.unitStride;
        const long long row = (long long)env * a.units + unit;
        float h[H];
#pragma unroll
        for (int o = 0; o < H; ++o) h[o] = b1[o];
        for (int k = 0; k < nIn; ++k) {
            const float xv = (float)xr[k];
            const float4 *wr = reinterpret_cast<const float4 *>(W1t + k * H);
#pragma unroll
            for (int o4 = 0; o4 < H / 4; ++o4) {
                const float4 wv = wr[o4];
                h[4 * o4] = fmaf(wv.x, xv, h[4 * o4]); h[4 * o4 + 1] = fmaf(wv.y, xv, h[4 * o4 + 1]);
                h[4 * o4 + 2] = fmaf(wv.z, xv, h[4 * o4 + 2]); h[4 * o4 + 3] = fmaf(wv.w, xv, h[4 * o4 + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < H; ++o) h[o] = tanhf(h[o]);
        float best = -INFINITY;
        int arg = 0;
        for (int o = 0; o < A; ++o) {
            float acc = b2[o];
            const float4 *wr = reinterpret_cast<const float4 *>(W2 + o * H);
#pragma unroll
            for (int k4 = 0; k4 < H / 4; ++k4) {
                const float4 wv = wr[k4];
                acc = fmaf(wv.x, h[4 * k4], acc); acc = fmaf(wv.y, h[4 * k4 + 1], acc);
                acc = fmaf(wv.z, h[4 * k4 + 2], acc); acc = fmaf(wv.w, h[4 * k4 + 3], acc);
            }
            if (q.qOut) q.qOut[(size_t)row * A + o] = acc;
            if (acc > best) { best = acc; arg = o; }
        }
        float u0, u1;
        if (a.uOverride) {
            u0 = a.uOverride[2 * row];
            u1 = a.uOverride[2 * row + 1];
        } else {
            const unsigned long long g = (unsigned long long)(a.rowOffset + row);
            const unsigned long long stp = a.stepDev ? *a.stepDev : a.step;
            uint32_t x4[4];
            philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stp,
                          (kStreamPolicy << 28) | (uint32_t)((stp >> 32) & 0x0fffffffu), (uint32_t)a.seed,
                          (uint32_t)(a.seed >> 32), x4);
            u0 = (float)(x4[0] >> 8) * (1.0f / 16777216.0f);
            u1 = (float)(x4[1] >> 8) * (1.0f / 16777216.0f);
        }
        // `sample > eps_treshold` exploits, otherwise random.randrange(numberOfActions)
        int act = arg;
        if (!(u0 > q.epsilon)) {
            act = (int)(u1 * (float)A);
            act = act >= A ? A - 1 : act;
        }
        if (a.action) a.action[row] = act;
        if (a.actionRec) a.actionRec[(size_t)env * a.actionRecStride + unit] = (int16_t)act;
    }
}

#endif  // MSCHED_ACTOR_DISPATCH_TU

}  // namespace msched
