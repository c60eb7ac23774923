// policy_step_tc_kernel.cuh -- msched_policy_step on the 5th-generation tensor cores.
//
// Same work and same contract as policy_step_kernel.cuh (every PPO unit of a rollout step in one launch:
// src/PPOmodules.py:32-39,53-63,114-125,312-332), but the three Linear layers of a 128-row tile are
// tcgen05.mma instructions with the accumulator in tensor memory; the threads only do what a matrix unit cannot:
// int16 -> float, bias + Tanh, the hi/lo operand split, Softmax / Categorical.sample / log_prob.
//
//   rows      a CTA of 128 threads serves ONE unit (an acceptor unit, or an offer unit = core chooser followed by
//             the price chooser) and walks over 128-environment tiles; thread r = row r of the tile
//   operands  shared memory, canonical no-swizzle K-major layout (8-row x 16-byte core matrices; a K chunk of 4
//             floats of all 128 rows is one 2 KB panel, thread r owns 16 bytes of it: conflict-free 128-bit stores)
//   layer 1   the inputs are small integers (|x| <= 2047: exactly representable in TF32), so A needs no split:
//             D = X * W1hi^T + X * W1lo^T, two MMAs per K step of 8
//   layers 2,3  "3xTF32": h = hi + lo, D = Hhi*Whi^T + Hhi*Wlo^T + Hlo*Whi^T (about 2^-21 relative per product,
//             the order of the fp32 accumulation itself)
//   scales    2*log2(e) folded into W1, b1, W2, b2 (tanh = 1 - 2/(2^z' + 1): EX2, FADD, RCP, FFMA) and log2(e)
//             into W3, b3 (base-2 logits for the softmax)
//   loads     the observation rows of a warp's 32 environments are read warp-cooperatively (8 lanes per 32-byte
//             row: whole sectors), converted and written straight into the layer-1 A panels; the same lanes
//             write the experience-buffer copy of the row
//   price chooser  its four inputs are picked out of the layer-1 A panels (already floats) by the sampled core
// One elected thread issues the MMAs of a layer and commits them to an mbarrier; the CTA meets at a barrier before
// every issue (operands written, previous accumulator read).  Several CTAs per SM (30 KB of shared memory, 32
// tensor-memory columns each) keep the SM busy while a CTA waits for its MMAs.
#pragma once
#include "policy_step_kernel.cuh"
#include "tc_primitives.cuh"

namespace msched {

// B operands (hi then lo) of one 16-wide net + its scaled biases, bytes
template <int KC1>  // K chunks (4 floats) of layer 1; even (UMMA K = 8 for tf32)
struct TcNetImage {
    static constexpr int kW1 = 0, kW1Half = KC1 * 256;       // [KC1 chunks][16 rows][16 B]
    static constexpr int kW2 = kW1 + 2 * kW1Half, kWHalf = 4 * 256;
    static constexpr int kW3 = kW2 + 2 * kWHalf;
    static constexpr int kBias = kW3 + 2 * kWHalf;            // b1 | b2 | b3, 16 floats each
    static constexpr int kBytes = kBias + 48 * 4;
    // all threads; W1 row position = lead + input index (the int16 position inside the row words)
    __device__ static void stage(unsigned char *s, const float *__restrict__ w, int nIn, int lead, int A)
    {
        constexpr float s2 = 2.f * kLog2e;
        const float *w2 = w + 16 * nIn + 16, *w3 = w2 + 256 + 16;
        for (int i = threadIdx.x; i < 16 * KC1 * 4; i += blockDim.x) {
            const int n = i / (KC1 * 4), pos = i - n * (KC1 * 4), k = pos - lead;
            const float v = (k >= 0 && k < nIn) ? w[n * nIn + k] * s2 : 0.f;
            const float hi = tf32_hi(v);
            const int off = (pos >> 2) * 256 + n * 16 + (pos & 3) * 4;
            *reinterpret_cast<float *>(s + kW1 + off) = hi;
            *reinterpret_cast<float *>(s + kW1 + kW1Half + off) = tf32_hi(v - hi);
        }
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            const int n = i >> 4, k = i & 15;
            const int off = (k >> 2) * 256 + n * 16 + (k & 3) * 4;
            const float v2 = w2[n * 16 + k] * s2, h2 = tf32_hi(v2);
            *reinterpret_cast<float *>(s + kW2 + off) = h2;
            *reinterpret_cast<float *>(s + kW2 + kWHalf + off) = tf32_hi(v2 - h2);
            const float v3 = n < A ? w3[n * 16 + k] * kLog2e : 0.f, h3 = tf32_hi(v3);
            *reinterpret_cast<float *>(s + kW3 + off) = h3;
            *reinterpret_cast<float *>(s + kW3 + kWHalf + off) = tf32_hi(v3 - h3);
        }
        float *b = reinterpret_cast<float *>(s + kBias);
        for (int i = threadIdx.x; i < 16; i += blockDim.x) {
            b[i] = w[16 * nIn + i] * s2;
            b[16 + i] = w2[256 + i] * s2;
            b[32 + i] = i < A ? w3[A * 16 + i] * kLog2e : -INFINITY;
        }
    }
};

// issue one layer: D[128 x 16] = A[128 x 4*KC] * W^T; aLo == 0: exact A (two MMAs per K step), else 3xTF32
__device__ __forceinline__ void tc_issue_layer(uint32_t tmemD, uint32_t aHi, uint32_t aLo, uint32_t bHi, uint32_t bLo, int KC,
                                               uint64_t *bar)
{
    constexpr uint32_t idesc = umma_idesc_tf32(128, 16);
    uint32_t acc = 0u;
    for (int ks = 0; ks < KC / 2; ++ks) {
        const uint64_t ah = umma_smem_desc(aHi + ks * 4096, 2048u, 128u);
        const uint64_t bh = umma_smem_desc(bHi + ks * 512, 256u, 128u);
        const uint64_t bl = umma_smem_desc(bLo + ks * 512, 256u, 128u);
        umma_tf32(tmemD, ah, bh, idesc, acc);
        umma_tf32(tmemD, ah, bl, idesc, 1u);
        if (aLo) umma_tf32(tmemD, umma_smem_desc(aLo + ks * 4096, 2048u, 128u), bh, idesc, 1u);
        acc = 1u;
    }
    umma_commit(bar);
}

// accumulator row + bias -> Tanh -> hi / lo A panels of the next layer (K = 16: 4 + 4 panels of 2 KB)
__device__ __forceinline__ void tc_hidden_epilogue(uint32_t trow, const float *__restrict__ bias, unsigned char *aH, int tid)
{
    float v[16];
    tmem_ld16(trow, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + 4 * c);
        const float h0 = tanh_scaled(v[4 * c] + b4.x), h1 = tanh_scaled(v[4 * c + 1] + b4.y);
        const float h2 = tanh_scaled(v[4 * c + 2] + b4.z), h3 = tanh_scaled(v[4 * c + 3] + b4.w);
        const float i0 = tf32_hi(h0), i1 = tf32_hi(h1), i2 = tf32_hi(h2), i3 = tf32_hi(h3);
        *reinterpret_cast<float4 *>(aH + c * 2048 + tid * 16) = make_float4(i0, i1, i2, i3);
        *reinterpret_cast<float4 *>(aH + (4 + c) * 2048 + tid * 16) = make_float4(h0 - i0, h1 - i1, h2 - i2, h3 - i3);
    }
}

// the barrier every MMA issue sits behind: operand stores visible to the async proxy, accumulator reads done
__device__ __forceinline__ void tc_cta_sync()
{
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// logits of the thread's row -> sample; AP = 8 or 16 columns are read
template <int AP>
__device__ __forceinline__ int tc_sample(uint32_t trow, const float *__restrict__ b3, int A, float u, float &logp, float *probsOut)
{
    float v[16];
    tmem_ld16(trow, v);
    float lg[AP];
#pragma unroll
    for (int o = 0; o < AP; ++o) lg[o] = v[o] + b3[o];
    return sample_row<AP>(lg, A, u, logp, probsOut);
}

// observation rows of the warp's 32 environments -> layer-1 A panels (floats, exact) and the experience buffer.
// LPR lanes per row read consecutive words; the word's two values land at K positions 2w, 2w+1 of the row
template <int KW>
__device__ __forceinline__ void tc_load_rows(const PolicyStepArgs &a, const PolicyGroupArgs &g, int tileEnv0, int unit, int offWords,
                                             unsigned char *aX)
{
    constexpr int LPR = KW <= 2 ? 2 : KW <= 4 ? 4 : KW <= 8 ? 8 : KW <= 16 ? 16 : 32, RPI = 32 / LPR;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    const int w = lane % LPR, rs = lane / LPR;
    const uint32_t *ob = reinterpret_cast<const uint32_t *>(a.obs);
    const long long strideW = a.obsStride >> 1;
    uint32_t *xu = reinterpret_cast<uint32_t *>(g.xUsed);
    const int xuW = g.xUsedStride >> 1;
#pragma unroll 4
    for (int r0 = 0; r0 < 32; r0 += RPI) {
        const int r = wbase + r0 + rs, env = tileEnv0 + r;
        if (w < KW) {
            const bool live = env < a.nEnvs;
            const uint32_t v = live ? ob[(size_t)env * strideW + offWords + w] : 0u;
            float lo, hi;
            halves_to_float(v, lo, hi);
            // K position 2w -> chunk w/2, element 2*(w&1); the pair is 8 aligned bytes of the row's 16-byte slot
            *reinterpret_cast<float2 *>(aX + (w >> 1) * 2048 + r * 16 + (w & 1) * 8) = make_float2(lo, hi);
            if (xu && live) xu[((size_t)env * g.units + unit) * xuW + w] = v;
        }
    }
}

// KW_A / KW_O: words per acceptor / offer row (even number of K chunks after padding); AP_*: logits columns read
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P>
struct PolicyStepTcSmem {
    static constexpr int KC_A = ((2 * KW_A + 7) / 8) * 2, KC_O = ((2 * KW_O + 7) / 8) * 2;  // layer-1 K chunks, even
    static constexpr int kNetA = TcNetImage<KC_A>::kBytes;
    static constexpr int kNetO = TcNetImage<KC_O>::kBytes + (AP_P > 0 ? TcNetImage<2>::kBytes : 0);
    static constexpr int kNet = ((kNetA > kNetO ? kNetA : kNetO) + 127) & ~127;
    static constexpr int KCX = KC_A > KC_O ? KC_A : KC_O;
    static constexpr int kAX = kNet, kAH = kAX + KCX * 2048;
    static constexpr int kBytes = kAH + 8 * 2048;
};

template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P>
__global__ void __launch_bounds__(128, 5) policy_step_tc_kernel(const __grid_constant__ PolicyStepArgs a)
{
    using SM = PolicyStepTcSmem<KW_A, AP_A, KW_O, AP_O, AP_P>;
    extern __shared__ __align__(128) unsigned char smc[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *aX = smc + SM::kAX, *aH = smc + SM::kAH;
    const int nAccCtas = a.acc.units * a.ctasPerAccUnit;
    const bool isAcc = (int)blockIdx.x < nAccCtas;
    const int nTiles = (a.nEnvs + 127) / 128;

    if (warp == 0) tmem_alloc(&tmemBase, 32u);
    if (tid == 32) mbar_init(&bar, 1);
    int unit, slice, stride;
    if (isAcc) {
        unit = blockIdx.x / a.ctasPerAccUnit; slice = blockIdx.x - unit * a.ctasPerAccUnit; stride = a.ctasPerAccUnit;
        const PolicyGroupArgs &g = a.acc;
        const int net = (unit / g.unitDiv) % g.nNets;
        const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
        TcNetImage<SM::KC_A>::stage(smc, g.weights + (size_t)net * pc, g.nIn, g.xOffset & 1, g.nActions);
    } else {
        const int id = blockIdx.x - nAccCtas;
        unit = id / a.ctasPerOffUnit; slice = id - unit * a.ctasPerOffUnit; stride = a.ctasPerOffUnit;
        const PolicyGroupArgs &g = a.core, &gp = a.price;
        const int net = (unit / g.unitDiv) % g.nNets;
        const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
        TcNetImage<SM::KC_O>::stage(smc, g.weights + (size_t)net * pc, g.nIn, 0, g.nActions);
        if constexpr (AP_P > 0) {
            const int netp = (unit / gp.unitDiv) % gp.nNets;
            const int pcp = 16 * 4 + 16 + 256 + 16 + 16 * gp.nActions + gp.nActions;
            TcNetImage<2>::stage(smc + TcNetImage<SM::KC_O>::kBytes, gp.weights + (size_t)netp * pcp, 4, 0, gp.nActions);
        }
    }
    // zero the A panels once: K positions beyond a row's words stay zero for the whole kernel
    for (int i = tid; i < (SM::KCX + 8) * 2048 / 16; i += 128) reinterpret_cast<uint4 *>(aX)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_cta_sync();
    const uint32_t tbase = tmemBase;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t sNet = smem_u32(smc), sAX = smem_u32(aX), sAH = smem_u32(aH);
    uint32_t ph = 0u;

    for (int tile = slice; tile < nTiles; tile += stride) {
        const int env = tile * 128 + tid;
        const bool live = env < a.nEnvs;
        if (isAcc) {
            using NI = TcNetImage<SM::KC_A>;
            const PolicyGroupArgs &g = a.acc;
            const float *bias = reinterpret_cast<const float *>(smc + NI::kBias);
            tc_load_rows<KW_A>(a, g, tile * 128, unit, (g.xOffset - (g.xOffset & 1) + unit * g.xStride) >> 1, aX);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAX, 0u, sNet + NI::kW1, sNet + NI::kW1 + NI::kW1Half, SM::KC_A, &bar);
            float u;
            if (g.uOverride) {
                u = live ? g.uOverride[(size_t)env * g.units + unit] : 0.f;
            } else {
                uint32_t r[4];
                pair_draws(a, g.seed, env, unit, r);
                u = u24((env & 1) ? r[1] : r[0]);
            }
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            tc_hidden_epilogue(trow, bias, aH, tid);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNet + NI::kW2, sNet + NI::kW2 + NI::kWHalf, 4, &bar);
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            tc_hidden_epilogue(trow, bias + 16, aH, tid);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNet + NI::kW3, sNet + NI::kW3 + NI::kWHalf, 4, &bar);
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            float lp;
            const int act = tc_sample<AP_A>(trow, bias + 32, g.nActions, u, lp,
                                            (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
            if (live) emit_row(a, g, env, unit, act, lp, act);
        } else {
            using NI = TcNetImage<SM::KC_O>;
            const PolicyGroupArgs &g = a.core, &gp = a.price;
            const float *bias = reinterpret_cast<const float *>(smc + NI::kBias);
            tc_load_rows<KW_O>(a, g, tile * 128, unit, (g.xOffset + unit * g.xStride) >> 1, aX);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAX, 0u, sNet + NI::kW1, sNet + NI::kW1 + NI::kW1Half, SM::KC_O, &bar);
            float u, v = 0.f;
            if (g.uOverride) {
                u = live ? g.uOverride[(size_t)env * g.units + unit] : 0.f;
                if (AP_P > 0 && gp.uOverride) v = live ? gp.uOverride[(size_t)env * gp.units + unit] : 0.f;
            } else {
                uint32_t r[4];
                pair_draws(a, g.seed, env, unit, r);
                u = u24((env & 1) ? r[1] : r[0]);
                v = u24((env & 1) ? r[3] : r[2]);
            }
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            tc_hidden_epilogue(trow, bias, aH, tid);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNet + NI::kW2, sNet + NI::kW2 + NI::kWHalf, 4, &bar);
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            tc_hidden_epilogue(trow, bias + 16, aH, tid);
            tc_cta_sync();
            if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNet + NI::kW3, sNet + NI::kW3 + NI::kWHalf, 4, &bar);
            mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
            float lp;
            const int c = tc_sample<AP_O>(trow, bias + 32, g.nActions, u, lp,
                                          (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
            if (live) emit_row(a, g, env, unit, c, lp, c);
            if constexpr (AP_P > 0) {
                // FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): [core prio, core rem, slot prio, slot rem] of
                // the chosen core = K positions 2c, 2c+1, 2*nCores, 2*nCores+1 of the thread's own layer-1 row (floats);
                // core action 0 feeds the dummy [-5,-5,-5,-5] and reports price -5 (quirk Q1)
                using NP = TcNetImage<2>;
                const unsigned char *netP = smc + NI::kBytes;
                const float *biasP = reinterpret_cast<const float *>(netP + NP::kBias);
                const bool dummy = c <= 0 || c > a.nCores;
                const int cc = dummy ? 0 : c;
                const float2 pc2 = *reinterpret_cast<const float2 *>(aX + (cc >> 1) * 2048 + tid * 16 + (cc & 1) * 8);
                const float2 ps2 = *reinterpret_cast<const float2 *>(aX + (a.nCores >> 1) * 2048 + tid * 16 + (a.nCores & 1) * 8);
                const float4 in = dummy ? make_float4(-5.f, -5.f, -5.f, -5.f) : make_float4(pc2.x, pc2.y, ps2.x, ps2.y);
                if (gp.xUsed && live)
                    *reinterpret_cast<short4 *>(gp.xUsed + ((size_t)env * gp.units + unit) * gp.xUsedStride) =
                        make_short4((short)in.x, (short)in.y, (short)in.z, (short)in.w);
                // (a thread only ever reads and rewrites its OWN 16-byte slots of the panels: no barrier needed here)
                *reinterpret_cast<float4 *>(aX + tid * 16) = in;
                *reinterpret_cast<float4 *>(aX + 2048 + tid * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
                tc_cta_sync();
                const uint32_t sNetP = sNet + NI::kBytes;
                if (tid == 0) tc_issue_layer(tbase, sAX, 0u, sNetP + NP::kW1, sNetP + NP::kW1 + NP::kW1Half, 2, &bar);
                mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
                tc_hidden_epilogue(trow, biasP, aH, tid);
                tc_cta_sync();
                if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNetP + NP::kW2, sNetP + NP::kW2 + NP::kWHalf, 4, &bar);
                mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
                tc_hidden_epilogue(trow, biasP + 16, aH, tid);
                tc_cta_sync();
                if (tid == 0) tc_issue_layer(tbase, sAH, sAH + 4 * 2048, sNetP + NP::kW3, sNetP + NP::kW3 + NP::kWHalf, 4, &bar);
                mbar_wait_bounded(&bar, ph); ph ^= 1u; tc_fence_after();
                float lq;
                const int b = tc_sample<AP_P>(trow, biasP + 32, gp.nActions, v, lq,
                                              (gp.probs && live) ? gp.probs + ((size_t)env * gp.units + unit) * gp.nActions : nullptr);
                if (live) emit_row(a, gp, env, unit, b, lq, c == 0 ? -5 : b);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 32u);
}

}  // namespace msched
