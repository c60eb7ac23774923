// policy_kernels.cuh -- the PPO pieces of the rollout path.
//
//   src/PPOmodules.py:32-39,53-63   ActorCritic.actor + act(): Linear-Tanh-Linear-Tanh-Linear-
//                                   Softmax, Categorical.sample, Categorical.log_prob
//   src/PPOmodules.py:114-125       PPO.selectAction (state.float())
//   src/PPOmodules.py:128-137       PPO.update returns prologue
#pragma once
#include "msched_common.cuh"

namespace msched {

// ---- discounted returns --------------------------------------------------------------------
// One lane per unit m; rewards/out are [T][M] so that a warp touches 128 contiguous bytes per
// time step.  G_t = r_t + gamma*G_{t+1} in float64 like the Python loop, cast to float32; the
// normalisation statistics are accumulated in float64 over the float32 values in the same sweep
// and a second sweep recomputes G instead of re-reading it (12 B/element of HBM traffic).
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
    const float denom = sd + 1e-7f;
    disc = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        disc = __dadd_rn((double)r[(size_t)t * M + m], __dmul_rn(gamma, disc));
        out[(size_t)t * M + m] = ((float)disc - mean) / denom;
    }
}

// ---- actor forward, fp32 SIMT version --------------------------------------------------------
// grid = (ceil(n_envs / 128), units); a CTA evaluates ONE unit (one net) for 128 consecutive
// environments, so every lane reads the same weight at the same time (shared-memory broadcast,
// 128-bit) and FFMA is the bound.  Weights are staged transposed ([in][out]) in shared memory.
struct ActorArgs {
    const float *weights;  // n_nets * param_count floats, torch layout per net
    const int16_t *x;
    long long envStride, unitStride;  // in int16 elements
    int nIn, nHidden, nActions, nNets, unitDiv, units, nEnvs;
    unsigned long long seed, step;
    long long rowOffset;
    const float *uOverride;  // [M] or null
    int32_t *action;         // [M] or null
    float *logprob;          // [M] or null
    float *probs;            // [M][A] or null
    int16_t *actionRec;      // reported action also stored at actionRec[env*actionRecStride + unit]
    long long actionRecStride;
    const int32_t *gatherCore;  // FreePriceOfferPPO price chooser: gather the 4 inputs by this action
    int16_t *xUsed;             // [M][nIn] input actually fed, or null
    int nCores;
};

constexpr int kActorMaxActions = 64;

template <int H>
__global__ void __launch_bounds__(128) actor_forward_simt(const ActorArgs a)
{
    extern __shared__ __align__(16) float sw[];
    const int nIn = a.nIn, A = a.nActions;
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + H * H + H + A * H + A;
    const float *w = a.weights + (size_t)net * pc;
    // smem layout: W1t [nIn][H] | b1 [H] | W2t [H][H] | b2 [H] | W3t [H][Apad] | b3 [Apad]
    const int Apad = (A + 3) & ~3;
    float *W1t = sw, *b1 = W1t + nIn * H, *W2t = b1 + H, *b2 = W2t + H * H, *W3t = b2 + H,
          *b3 = W3t + H * Apad;
    for (int i = threadIdx.x; i < H * nIn; i += blockDim.x) W1t[(i % nIn) * H + i / nIn] = w[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = w[H * nIn + i];
    const float *w2 = w + H * nIn + H;
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) W2t[(i % H) * H + i / H] = w2[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b2[i] = w2[H * H + i];
    const float *w3 = w2 + H * H + H;
    for (int i = threadIdx.x; i < H * Apad; i += blockDim.x) W3t[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < A * H; i += blockDim.x) W3t[(i % H) * Apad + i / H] = w3[i];
    for (int i = threadIdx.x; i < Apad; i += blockDim.x) b3[i] = i < A ? w3[A * H + i] : 0.f;
    __syncthreads();

    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= a.nEnvs) return;
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

    float h1[H], h2[H];
#pragma unroll
    for (int o = 0; o < H; ++o) h1[o] = b1[o];
    for (int k = 0; k < nIn; ++k) {
        const float xv = a.gatherCore ? (float)g4[k & 3] : (float)xr[k];
        const float4 *wr = reinterpret_cast<const float4 *>(W1t + k * H);
#pragma unroll
        for (int o4 = 0; o4 < H / 4; ++o4) {
            const float4 wv = wr[o4];
            h1[4 * o4 + 0] = fmaf(wv.x, xv, h1[4 * o4 + 0]);
            h1[4 * o4 + 1] = fmaf(wv.y, xv, h1[4 * o4 + 1]);
            h1[4 * o4 + 2] = fmaf(wv.z, xv, h1[4 * o4 + 2]);
            h1[4 * o4 + 3] = fmaf(wv.w, xv, h1[4 * o4 + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < H; ++o) { h1[o] = tanhf(h1[o]); h2[o] = b2[o]; }
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
    for (int o = 0; o < H; ++o) h2[o] = tanhf(h2[o]);

    float lg[kActorMaxActions];
#pragma unroll
    for (int o = 0; o < kActorMaxActions; ++o) lg[o] = 0.f;
    float mx = -INFINITY;
#pragma unroll
    for (int o4 = 0; o4 < kActorMaxActions / 4; ++o4) {
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
#pragma unroll
    for (int o = 0; o < kActorMaxActions; ++o)
        if (o < A) mx = fmaxf(mx, lg[o]);
    float sum = 0.f;
#pragma unroll
    for (int o = 0; o < kActorMaxActions; ++o)
        if (o < A) { lg[o] = expf(lg[o] - mx); sum += lg[o]; }
    float tot = 0.f;  // Categorical(probs) renormalises by the sum of the softmax output
#pragma unroll
    for (int o = 0; o < kActorMaxActions; ++o)
        if (o < A) { lg[o] = lg[o] / sum; tot += lg[o]; }
    if (a.probs)
        for (int o = 0; o < A; ++o) a.probs[(size_t)row * A + o] = lg[o];
    if (!a.action && !a.logprob && !a.actionRec) return;
    float u;
    if (a.uOverride) {
        u = a.uOverride[row];
    } else {
        const unsigned long long g = (unsigned long long)(a.rowOffset + row);
        uint32_t x4[4];
        philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)a.step,
                      (kStreamPolicy << 28) | (uint32_t)((a.step >> 32) & 0x0fffffffu), (uint32_t)a.seed,
                      (uint32_t)(a.seed >> 32), x4);
        u = (float)(x4[0] >> 8) * (1.0f / 16777216.0f);
    }
    const float thr = u * tot;
    float cdf = 0.f, pa = 0.f;
    int act = A - 1;
    bool found = false;
#pragma unroll
    for (int o = 0; o < kActorMaxActions; ++o)
        if (o < A) {
            cdf += lg[o];
            if (!found && cdf > thr) { act = o; pa = lg[o]; found = true; }
        }
    if (!found) {
#pragma unroll
        for (int o = 0; o < kActorMaxActions; ++o)
            if (o == A - 1) pa = lg[o];
    }
    if (a.action) a.action[row] = act;  // what PPO.selectAction stores in buffer.actions
    if (a.actionRec)                    // what the world is handed (price -5 next to core action 0)
        a.actionRec[(size_t)env * a.actionRecStride + unit] = (int16_t)((a.gatherCore && gsel == 0) ? -5 : act);
    if (a.logprob) {
        const float eps = 1.1920928955078125e-07f;
        float pn = pa / tot;
        pn = fminf(fmaxf(pn, eps), 1.f - eps);
        a.logprob[row] = logf(pn);
    }
}

inline int launch_actor_forward(const MschedMlpGroup &g, const MschedActorIO &io, cudaStream_t s)
{
    if (g.n_actions > kActorMaxActions) return -1;
    ActorArgs a;
    a.weights = g.weights; a.x = io.x;
    a.envStride = io.env_stride ? io.env_stride : (long long)io.x_stride * io.units; a.unitStride = io.x_stride;
    a.nIn = g.n_in; a.nHidden = g.n_hidden; a.nActions = g.n_actions; a.nNets = g.n_nets;
    a.unitDiv = g.unit_div > 0 ? g.unit_div : 1;
    a.units = io.units; a.nEnvs = io.n_envs;
    a.seed = io.seed; a.step = io.step; a.rowOffset = io.row_offset; a.uOverride = io.u_override;
    a.action = io.action; a.logprob = io.logprob; a.probs = io.probs;
    a.actionRec = io.action_rec; a.actionRecStride = io.action_rec_stride;
    a.gatherCore = io.gather_core; a.xUsed = io.x_used; a.nCores = io.n_cores;
    const int H = g.n_hidden, Apad = (g.n_actions + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)g.n_in * H + H + (size_t)H * H + H + (size_t)H * Apad + Apad);
    dim3 grid((a.nEnvs + 127) / 128, io.units);
    if (H == 16) actor_forward_simt<16><<<grid, 128, smem, s>>>(a);
    else if (H == 32) actor_forward_simt<32><<<grid, 128, smem, s>>>(a);
    else if (H == 64) actor_forward_simt<64><<<grid, 128, smem, s>>>(a);
    else return -1;
    return 0;
}

}  // namespace msched
