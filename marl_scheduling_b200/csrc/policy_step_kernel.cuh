// policy_step_kernel.cuh -- every PPO unit of one rollout step in ONE launch.
//
//   src/PPOmodules.py:32-39,53-63    ActorCritic.actor + act(): Linear-Tanh-Linear-Tanh-Linear-Softmax,
//                                    Categorical.sample, Categorical.log_prob
//   src/PPOmodules.py:114-125        PPO.selectAction: state / action / log-prob go to the experience buffer
//   src/PPOmodules.py:312-332        FreePriceOfferPPO.selectAction: core chooser, then the price chooser on
//                                    [core prio, core rem, slot prio, slot rem] of the chosen core (quirk Q1)
//   src/Agent.py:504-515, 589-601    the agents' getActions: every acceptor unit and every offer unit of a step
//
// The per-group kernels (policy_kernels.cuh, actor_tc_kernel.cuh) cost one launch per net group and an extra
// pass over the observations per group; at 65,536 environments the three launches of the free-price agents
// were 80 % of a rollout step.  Here one persistent grid covers all groups: a CTA serves ONE unit (= one net,
// staged once in shared memory, transposed, the Tanh scale 2*log2(e) and the Softmax scale log2(e) folded into
// the weights) and walks over 256-environment tiles.  What makes it cheaper per row than the per-group kernel:
//   * TWO rows (adjacent environments) per thread: every 128-bit weight load feeds 8 FFMAs instead of 4;
//   * the observation row is read as aligned 32-bit words straight into registers (16 independent loads in
//     flight per thread), int16 -> float on the integer / FMA pipes (bias trick, no I2F on the XU pipe that
//     the Tanh / Softmax exponentials keep busy);
//   * tanh = 1 - 2/(2^z' + 1) with z' already scaled: EX2, FADD, RCP, FFMA;
//   * one Philox call per environment PAIR and unit serves both rows and both choosers of an offer unit;
//   * an offer unit's price chooser runs in the thread that sampled the core: its 4 inputs are picked from
//     the row words already in registers;
//   * actions go straight into the environment's action record, action / log-prob (and, if asked, the input
//     rows) into the experience buffer slot of the step.
// This is the fp32 SIMT form.  It serves observations beyond +-511 (msched_policy_step's input_bound); within that
// bound the tensor-core kernel of policy_step_tc_kernel.cuh (same contract, same draws) runs instead: 47 us against 78 us
// at 65,536 environments of the config-3 shape.  Shared by both: the argument structs, sample_row, pair_draws, emit_row.
#pragma once
#include "msched_common.cuh"
#include "policy_common.cuh"

namespace msched {

constexpr uint32_t kStreamPolicyStep = 4;

struct PolicyGroupArgs {
    const float *weights;  // n_nets * param_count, torch layout
    int nIn, nActions, nNets, unitDiv, units;
    int xOffset, xStride;  // int16 offsets inside the observation record: row of unit u at xOffset + u * xStride
    int recOffset;         // int16 offset inside the action record: action of unit u at recOffset + u
    unsigned long long seed;
    int32_t *action;       // [nEnvs][units] or null
    float *logprob;        // [nEnvs][units] or null
    int16_t *xUsed;        // [nEnvs][units][xUsedStride] or null
    int xUsedStride;
    const float *uOverride;  // [nEnvs][units] or null
    float *probs;            // [nEnvs][units][nActions] or null (tests)
};

struct PolicyStepArgs {
    const int16_t *obs;
    long long obsStride;  // int16 per env
    int nEnvs, nCores;
    int16_t *actionRec;
    long long actionRecStride;
    long long envOffset;
    unsigned long long step;
    const unsigned long long *stepDev;
    PolicyGroupArgs acc, core, price;  // price.weights == null: fixed prices, the offer unit is the core chooser alone
    int ctasPerAccUnit, ctasPerOffUnit;
};

// shared-memory image of one 16-wide net: W1t [2*KW][16] (one row per int16 position of the observation row words,
// zero rows for the leading pad / trailing unused positions) | b1 [16] | W2t [16][16] | b2 [16] | W3t [16][AP] | b3 [AP]
template <int KW, int AP>
struct NetImage {
    static constexpr int kW1 = 0, kB1 = 2 * KW * 16, kW2 = kB1 + 16, kB2 = kW2 + 256, kW3 = kB2 + 16, kB3 = kW3 + 16 * AP;
    static constexpr int kFloats = (kB3 + AP + 3) & ~3;
    // all threads of the CTA; `w` = the net's parameters in torch layout [W1 16*nIn | b1 | W2 | b2 | W3 A*16 | b3];
    // input k of the net sits at int16 position lead + k of the row
    __device__ static void stage(float *s, const float *__restrict__ w, int nIn, int lead, int A)
    {
        constexpr float s2 = 2.f * kLog2e;
        const float *w2 = w + 16 * nIn + 16, *w3 = w2 + 256 + 16;
        for (int i = threadIdx.x; i < 2 * KW * 16; i += blockDim.x) {
            const int pos = i >> 4, o = i & 15, k = pos - lead;
            s[kW1 + i] = (k >= 0 && k < nIn) ? w[o * nIn + k] * s2 : 0.f;
        }
        for (int i = threadIdx.x; i < 256; i += blockDim.x) { const int k = i >> 4, o = i & 15; s[kW2 + i] = w2[o * 16 + k] * s2; }
        for (int i = threadIdx.x; i < 16; i += blockDim.x) { s[kB1 + i] = w[16 * nIn + i] * s2; s[kB2 + i] = w2[256 + i] * s2; }
        for (int i = threadIdx.x; i < 16 * AP; i += blockDim.x) {
            const int k = i / AP, o = i - k * AP;
            s[kW3 + i] = o < A ? w3[o * 16 + k] * kLog2e : 0.f;
        }
        for (int i = threadIdx.x; i < AP; i += blockDim.x) s[kB3 + i] = i < A ? w3[A * 16 + i] * kLog2e : -INFINITY;
    }
};

// the two int16 halves of an observation word as floats, on the integer / FMA pipes: flip the sign bits
// (h + 32768 as unsigned), plant each half in the mantissa of 2^23 * 1.5, subtract 2^23 * 1.5 + 32768
__device__ __forceinline__ void halves_to_float(uint32_t w, float &lo, float &hi)
{
    const uint32_t b = w ^ 0x80008000u;
    lo = __uint_as_float(__byte_perm(b, 0x4b400000u, 0x7610)) - 12615680.f;
    hi = __uint_as_float(__byte_perm(b, 0x4b400000u, 0x7632)) - 12615680.f;
}

__device__ __forceinline__ float tanh_scaled(float z2)  // tanh(z) given z2 = 2 * log2(e) * z
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ex2_approx(z2) + 1.f));
    return fmaf(-2.f, r, 1.f);
}

// acc[r][o] += W[k][o] * x[r] for the two rows, one 16-float weight row = four 128-bit broadcast loads
__device__ __forceinline__ void fma_row16(const float *__restrict__ wrow, float x0, float x1, float (&a0)[16], float (&a1)[16])
{
#pragma unroll
    for (int o4 = 0; o4 < 4; ++o4) {
        const float4 wv = *reinterpret_cast<const float4 *>(wrow + 4 * o4);
        a0[4 * o4 + 0] = fmaf(wv.x, x0, a0[4 * o4 + 0]); a1[4 * o4 + 0] = fmaf(wv.x, x1, a1[4 * o4 + 0]);
        a0[4 * o4 + 1] = fmaf(wv.y, x0, a0[4 * o4 + 1]); a1[4 * o4 + 1] = fmaf(wv.y, x1, a1[4 * o4 + 1]);
        a0[4 * o4 + 2] = fmaf(wv.z, x0, a0[4 * o4 + 2]); a1[4 * o4 + 2] = fmaf(wv.z, x1, a1[4 * o4 + 2]);
        a0[4 * o4 + 3] = fmaf(wv.w, x0, a0[4 * o4 + 3]); a1[4 * o4 + 3] = fmaf(wv.w, x1, a1[4 * o4 + 3]);
    }
}

// Per-thread vectors live in shared memory as planes of kPlane words (one word per thread, conflict free): the
// row words x [2*KW planes: word w of the even / odd row] and the activations h [32 planes].  That keeps the
// layer loops ROLLED (a dynamic index into a register array would go to local memory): the whole kernel's inner
// loops are a few hundred instructions and stay in the instruction cache, and only the 32 accumulators of the
// layer being computed sit in registers.
constexpr int kPlane = 129;

// the three layers for the thread's two rows: row words in planes sx[0 .. 2*nWords), activations through sh
template <int KW, int AP>
__device__ __forceinline__ void mlp_pair(const float *__restrict__ s, const uint32_t *__restrict__ sx, float *__restrict__ sh,
                                         int nWords, float (&lg0)[AP], float (&lg1)[AP])
{
    using NI = NetImage<KW, AP>;
    const int tid = threadIdx.x;
    {
        float h0[16], h1[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) { h0[o] = s[NI::kB1 + o]; h1[o] = h0[o]; }
#pragma unroll 2
        for (int w = 0; w < nWords; ++w) {
            float a0, b0, a1, b1;
            halves_to_float(sx[(2 * w) * kPlane + tid], a0, b0);
            halves_to_float(sx[(2 * w + 1) * kPlane + tid], a1, b1);
            fma_row16(s + NI::kW1 + (2 * w) * 16, a0, a1, h0, h1);
            fma_row16(s + NI::kW1 + (2 * w + 1) * 16, b0, b1, h0, h1);
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) { sh[(2 * o) * kPlane + tid] = tanh_scaled(h0[o]); sh[(2 * o + 1) * kPlane + tid] = tanh_scaled(h1[o]); }
    }
    {
        float g0[16], g1[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) { g0[o] = s[NI::kB2 + o]; g1[o] = g0[o]; }
#pragma unroll 4
        for (int k = 0; k < 16; ++k) fma_row16(s + NI::kW2 + k * 16, sh[(2 * k) * kPlane + tid], sh[(2 * k + 1) * kPlane + tid], g0, g1);
#pragma unroll
        for (int o = 0; o < 16; ++o) { sh[(2 * o) * kPlane + tid] = tanh_scaled(g0[o]); sh[(2 * o + 1) * kPlane + tid] = tanh_scaled(g1[o]); }
    }
#pragma unroll
    for (int o = 0; o < AP; ++o) { lg0[o] = s[NI::kB3 + o]; lg1[o] = lg0[o]; }
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const float y0 = sh[(2 * k) * kPlane + tid], y1 = sh[(2 * k + 1) * kPlane + tid];
#pragma unroll
        for (int o4 = 0; o4 < AP / 4; ++o4) {
            const float4 wv = *reinterpret_cast<const float4 *>(s + NI::kW3 + k * AP + 4 * o4);
            lg0[4 * o4] = fmaf(wv.x, y0, lg0[4 * o4]); lg0[4 * o4 + 1] = fmaf(wv.y, y0, lg0[4 * o4 + 1]);
            lg0[4 * o4 + 2] = fmaf(wv.z, y0, lg0[4 * o4 + 2]); lg0[4 * o4 + 3] = fmaf(wv.w, y0, lg0[4 * o4 + 3]);
            lg1[4 * o4] = fmaf(wv.x, y1, lg1[4 * o4]); lg1[4 * o4 + 1] = fmaf(wv.y, y1, lg1[4 * o4 + 1]);
            lg1[4 * o4 + 2] = fmaf(wv.z, y1, lg1[4 * o4 + 2]); lg1[4 * o4 + 3] = fmaf(wv.w, y1, lg1[4 * o4 + 3]);
        }
    }
}

// Softmax -> Categorical(probs): inverse-CDF sample with draw u, Categorical.log_prob semantics (renormalised,
// clamped to [eps, 1-eps]); lg are base-2 logits, -inf beyond the net's A actions
// AE = the number of columns that can hold an action (A <= AE <= AP; a kernel built for one net shape passes its exact
// action count and skips the padding columns' exponentials and sums)
template <int AP, int AE = AP>
__device__ __forceinline__ int sample_row(float (&lg)[AP], int A, float u, float &logp, float *probsOut)
{
    float mx = lg[0];
#pragma unroll
    for (int o = 1; o < AE; ++o) mx = fmaxf(mx, lg[o]);
    float sum = 0.f;
#pragma unroll
    for (int o = 0; o < AE; ++o) { lg[o] = ex2_approx(lg[o] - mx); sum += lg[o]; }
    const float inv = __fdividef(1.f, sum);
    float tot = 0.f;
#pragma unroll
    for (int o = 0; o < AE; ++o) { lg[o] *= inv; tot += lg[o]; }
    if (probsOut) {
        for (int o = 0; o < A; ++o) {
            float v = lg[0];
#pragma unroll
            for (int q = 1; q < AE; ++q) v = (q == o) ? lg[q] : v;
            probsOut[o] = v;
        }
    }
    // first action whose running sum exceeds u * tot, the last action if none does: the sums never decrease
    // (probabilities >= 0), so that index = the number of sums that do NOT exceed the threshold
    const float thr = u * tot;
    float cdf = 0.f;
    int act = 0;
#pragma unroll
    for (int o = 0; o < AE; ++o) {
        cdf += lg[o];
        act += (cdf > thr) ? 0 : 1;
    }
    act = min(act, A - 1);
    float pa = lg[0];
#pragma unroll
    for (int o = 1; o < AE; ++o) pa = (o == act) ? lg[o] : pa;
    const float eps = 1.1920928955078125e-07f;
    float pn = __fdividef(pa, tot);
    pn = fminf(fmaxf(pn, eps), 1.f - eps);
    logp = __logf(pn);
    return act;
}

__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// draws of an environment pair for one unit: words 0,1 = the two rows of the acceptor / core chooser, words 2,3 =
// the two rows of the price chooser.  Counter (pair lo, pair hi, step lo, 4 << 28 | step hi : 12 | unit : 16), key = seed
__device__ __forceinline__ unsigned long long policy_step_now(const PolicyStepArgs &a) { return a.stepDev ? *a.stepDev : a.step; }
__device__ __forceinline__ void pair_draws(const PolicyStepArgs &a, unsigned long long seed, int envLocal, int unit, uint32_t (&x)[4],
                                           unsigned long long stp)
{
    const unsigned long long pair = (unsigned long long)(a.envOffset + envLocal) >> 1;
    philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)stp,
                  (kStreamPolicyStep << 28) | ((uint32_t)((stp >> 32) & 0xfffu) << 16) | (uint32_t)(unit & 0xffff),
                  (uint32_t)seed, (uint32_t)(seed >> 32), x);
}

// The observation rows of the warp's 64 environments -> the x planes, warp-cooperatively: LPR lanes per row read the
// row's consecutive words (whole 32-byte sectors instead of 32 lanes x 4 bytes of 32 different rows)
template <int KW>
__device__ __forceinline__ void coop_load_rows(const PolicyStepArgs &a, int envBase, int offWords, uint32_t *sx)
{
    constexpr int LPR = KW <= 2 ? 2 : KW <= 4 ? 4 : KW <= 8 ? 8 : KW <= 16 ? 16 : 32, RPI = 32 / LPR;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    const int w = lane % LPR, rs = lane / LPR;
    const uint32_t *ob = reinterpret_cast<const uint32_t *>(a.obs);
    const long long strideW = a.obsStride >> 1;
#pragma unroll 4
    for (int r0 = 0; r0 < 64; r0 += RPI) {
        const int r = r0 + rs, env = envBase + r;
        if (w < KW) {
            const uint32_t v = env < a.nEnvs ? ob[(size_t)env * strideW + offWords + w] : 0u;
            sx[(2 * w + (r & 1)) * kPlane + wbase + (r >> 1)] = v;
        }
    }
}

// the x planes -> the experience buffer rows [env][unit][strideW words], the same lane mapping (full sectors)
template <int KW>
__device__ __forceinline__ void coop_store_rows(const PolicyStepArgs &a, const PolicyGroupArgs &g, int envBase, int unit, const uint32_t *sx)
{
    constexpr int LPR = KW <= 2 ? 2 : KW <= 4 ? 4 : KW <= 8 ? 8 : KW <= 16 ? 16 : 32, RPI = 32 / LPR;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    const int w = lane % LPR, rs = lane / LPR;
    uint32_t *dst = reinterpret_cast<uint32_t *>(g.xUsed);
    const int strideW = g.xUsedStride >> 1;
#pragma unroll 4
    for (int r0 = 0; r0 < 64; r0 += RPI) {
        const int r = r0 + rs, env = envBase + r;
        if (w < KW && env < a.nEnvs)
            dst[((size_t)env * g.units + unit) * strideW + w] = sx[(2 * w + (r & 1)) * kPlane + wbase + (r >> 1)];
    }
}

// outputs of one row
__device__ __forceinline__ void emit_row(const PolicyStepArgs &a, const PolicyGroupArgs &g, int env, int unit, int act, float logp,
                                         int reported)
{
    const size_t row = (size_t)env * g.units + unit;
    if (g.action) g.action[row] = act;
    if (g.logprob) g.logprob[row] = logp;
    if (a.actionRec) a.actionRec[(size_t)env * a.actionRecStride + g.recOffset + unit] = (int16_t)reported;
}

// KW_A / AP_A: words per acceptor row, padded action count; KW_O / AP_O: offer rows (core chooser); AP_P: price
// chooser (0 = none).  Shared memory: the unit's net image(s) | x planes | h planes.
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P>
struct PolicyStepSmem {
    static constexpr int APP = AP_P > 0 ? AP_P : 4;
    static constexpr int kImgA = NetImage<KW_A, AP_A>::kFloats;
    static constexpr int kImgO = NetImage<KW_O, AP_O>::kFloats + (AP_P > 0 ? NetImage<2, APP>::kFloats : 0);
    static constexpr int kImg = kImgA > kImgO ? kImgA : kImgO;
    static constexpr int kXPlanes = 2 * (KW_A > KW_O + 2 ? KW_A : KW_O + 2);  // the price chooser's 2 words sit behind the offer row
    static constexpr int kWords = kImg + (kXPlanes + 32) * kPlane;
};

template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P>
__global__ void __launch_bounds__(128, 5) policy_step_kernel(const __grid_constant__ PolicyStepArgs a)
{
    using SM = PolicyStepSmem<KW_A, AP_A, KW_O, AP_O, AP_P>;
    extern __shared__ __align__(16) float sw[];
    uint32_t *sx = reinterpret_cast<uint32_t *>(sw + SM::kImg);
    float *sh = sw + SM::kImg + SM::kXPlanes * kPlane;
    const int tid = threadIdx.x, wbase = tid & ~31;
    const int nAccCtas = a.acc.units * a.ctasPerAccUnit;
    const bool isAcc = (int)blockIdx.x < nAccCtas;
    const int nTiles = (a.nEnvs + 255) / 256;
    if (isAcc) {
        const int unit = blockIdx.x / a.ctasPerAccUnit, slice = blockIdx.x - unit * a.ctasPerAccUnit;
        const PolicyGroupArgs &g = a.acc;
        const int net = (unit / g.unitDiv) % g.nNets;
        const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
        const int lead = g.xOffset & 1;
        NetImage<KW_A, AP_A>::stage(sw, g.weights + (size_t)net * pc, g.nIn, lead, g.nActions);
        __syncthreads();
        const int offW = (g.xOffset - lead + unit * g.xStride) >> 1;
        for (int tile = slice; tile < nTiles; tile += a.ctasPerAccUnit) {
            const int envW = tile * 256 + 2 * wbase;  // first environment of this warp
            const int e0 = tile * 256 + 2 * tid, e1 = e0 + 1;
            const bool l0 = e0 < a.nEnvs, l1 = e1 < a.nEnvs;
            __syncwarp();
            coop_load_rows<KW_A>(a, envW, offW, sx);
            float u0, u1;
            if (g.uOverride) {
                u0 = l0 ? g.uOverride[(size_t)e0 * g.units + unit] : 0.f;
                u1 = l1 ? g.uOverride[(size_t)e1 * g.units + unit] : 0.f;
            } else {
                uint32_t r[4];
                pair_draws(a, g.seed, e0, unit, r, policy_step_now(a));
                u0 = u24(r[0]); u1 = u24(r[1]);
            }
            __syncwarp();
            float lg0[AP_A], lg1[AP_A];
            mlp_pair<KW_A, AP_A>(sw, sx, sh, KW_A, lg0, lg1);
            float lp0, lp1;
            const int a0 = sample_row<AP_A>(lg0, g.nActions, u0, lp0, (g.probs && l0) ? g.probs + ((size_t)e0 * g.units + unit) * g.nActions : nullptr);
            const int a1 = sample_row<AP_A>(lg1, g.nActions, u1, lp1, (g.probs && l1) ? g.probs + ((size_t)e1 * g.units + unit) * g.nActions : nullptr);
            if (l0) emit_row(a, g, e0, unit, a0, lp0, a0);
            if (l1) emit_row(a, g, e1, unit, a1, lp1, a1);
            if (g.xUsed) coop_store_rows<KW_A>(a, g, envW, unit, sx);
        }
    } else {
        const int id = blockIdx.x - nAccCtas;
        const int unit = id / a.ctasPerOffUnit, slice = id - unit * a.ctasPerOffUnit;
        const PolicyGroupArgs &g = a.core, &gp = a.price;
        constexpr int APP = SM::APP;
        float *swp = sw + NetImage<KW_O, AP_O>::kFloats;
        {
            const int net = (unit / g.unitDiv) % g.nNets;
            const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
            NetImage<KW_O, AP_O>::stage(sw, g.weights + (size_t)net * pc, g.nIn, 0, g.nActions);
            if (AP_P > 0) {
                const int netp = (unit / gp.unitDiv) % gp.nNets;
                const int pcp = 16 * 4 + 16 + 256 + 16 + 16 * gp.nActions + gp.nActions;
                NetImage<2, APP>::stage(swp, gp.weights + (size_t)netp * pcp, 4, 0, gp.nActions);
            }
        }
        __syncthreads();
        const int offW = (g.xOffset + unit * g.xStride) >> 1;
        uint32_t *sxp = sx + 2 * KW_O * kPlane;  // the price chooser's two input words per row
        for (int tile = slice; tile < nTiles; tile += a.ctasPerOffUnit) {
            const int envW = tile * 256 + 2 * wbase;
            const int e0 = tile * 256 + 2 * tid, e1 = e0 + 1;
            const bool l0 = e0 < a.nEnvs, l1 = e1 < a.nEnvs;
            __syncwarp();
            coop_load_rows<KW_O>(a, envW, offW, sx);
            float u0, u1, v0 = 0.f, v1 = 0.f;
            if (g.uOverride) {
                u0 = l0 ? g.uOverride[(size_t)e0 * g.units + unit] : 0.f;
                u1 = l1 ? g.uOverride[(size_t)e1 * g.units + unit] : 0.f;
                if (AP_P > 0 && gp.uOverride) {
                    v0 = l0 ? gp.uOverride[(size_t)e0 * gp.units + unit] : 0.f;
                    v1 = l1 ? gp.uOverride[(size_t)e1 * gp.units + unit] : 0.f;
                }
            } else {
                uint32_t r[4];
                pair_draws(a, g.seed, e0, unit, r, policy_step_now(a));
                u0 = u24(r[0]); u1 = u24(r[1]); v0 = u24(r[2]); v1 = u24(r[3]);
            }
            __syncwarp();
            int c0, c1;
            {
                float lg0[AP_O], lg1[AP_O];
                mlp_pair<KW_O, AP_O>(sw, sx, sh, KW_O, lg0, lg1);
                float lp0, lp1;
                c0 = sample_row<AP_O>(lg0, g.nActions, u0, lp0, (g.probs && l0) ? g.probs + ((size_t)e0 * g.units + unit) * g.nActions : nullptr);
                c1 = sample_row<AP_O>(lg1, g.nActions, u1, lp1, (g.probs && l1) ? g.probs + ((size_t)e1 * g.units + unit) * g.nActions : nullptr);
                if (l0) emit_row(a, g, e0, unit, c0, lp0, c0);
                if (l1) emit_row(a, g, e1, unit, c1, lp1, c1);
            }
            if (g.xUsed) coop_store_rows<KW_O>(a, g, envW, unit, sx);
            if (AP_P > 0) {
                // FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): the price net sees [core prio, core rem,
                // slot prio, slot rem] of the chosen core = row words c and nCores; core action 0 feeds the dummy
                // [-5,-5,-5,-5] and reports price -5 (quirk Q1)
                const uint32_t dummy = 0xfffbfffbu;
                const bool d0 = c0 <= 0 || c0 > a.nCores, d1 = c1 <= 0 || c1 > a.nCores;
                sxp[0 * kPlane + tid] = d0 ? dummy : sx[(2 * c0) * kPlane + tid];
                sxp[1 * kPlane + tid] = d1 ? dummy : sx[(2 * c1 + 1) * kPlane + tid];
                sxp[2 * kPlane + tid] = d0 ? dummy : sx[(2 * a.nCores) * kPlane + tid];
                sxp[3 * kPlane + tid] = d1 ? dummy : sx[(2 * a.nCores + 1) * kPlane + tid];
                float q0[APP], q1[APP];
                mlp_pair<2, APP>(swp, sxp, sh, 2, q0, q1);
                float lq0, lq1;
                const int b0 = sample_row<APP>(q0, gp.nActions, v0, lq0, (gp.probs && l0) ? gp.probs + ((size_t)e0 * gp.units + unit) * gp.nActions : nullptr);
                const int b1 = sample_row<APP>(q1, gp.nActions, v1, lq1, (gp.probs && l1) ? gp.probs + ((size_t)e1 * gp.units + unit) * gp.nActions : nullptr);
                if (l0) emit_row(a, gp, e0, unit, b0, lq0, c0 == 0 ? -5 : b0);
                if (l1) emit_row(a, gp, e1, unit, b1, lq1, c1 == 0 ? -5 : b1);
                __syncwarp();
                if (gp.xUsed) coop_store_rows<2>(a, gp, envW, unit, sxp);
            }
        }
    }
}

}  // namespace msched
