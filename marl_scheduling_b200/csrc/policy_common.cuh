// policy_common.cuh -- argument record and math helpers shared by the policy-side kernels (actor forward on
// tcgen05 / warp-level MMA / fp32 SIMT, PPO.update gradient); templates and inline device functions only, so
// that every translation unit of libmsched.so may include it.
#pragma once
#include "msched_common.cuh"

namespace msched {

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
    unsigned long long *timeline;  // diagnostics, or null
    const unsigned long long *stepDev;  // device-side step counter (CUDA-graph replays), or null -> `step`
};

constexpr int kActorMaxActions = 64;

// Softmax -> Categorical(probs): sample by inverse CDF, log_prob with torch's renormalisation and
// clamp to [eps, 1-eps] (src/PPOmodules.py:53-63); one row per thread, logits in registers
// tanh with 2 MUFU ops: 1 - 2/(e^{2x}+1).  Absolute error < 3e-7 on the whole range (what matters
// downstream: the activations feed a Linear layer), exact limits for |x| -> inf, tanh(0) = 0
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_tanh(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));  // e^{2x}
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

// ---- warp-level tensor-core helpers: TF32 with error compensation ("3xTF32") ----
// x = hi + lo with both parts representable in TF32; a product is hi*hi + hi*lo + lo*hi (about 2^-21 relative)
// The split is a mantissa mask, not cvt.rna.tf32: the conversion instruction runs on the XU pipe (the
// transcendental unit, a quarter-rate pipe that tanh and exp already keep busy -- ncu showed it 89 % active in
// the first PPO gradient kernel), the mask and the subtraction run on the ALU / FMA pipes.  hi = v with the low
// 13 mantissa bits cleared (exactly representable), lo = v - hi (exact) cut to TF32 the same way:
// |v - hi - lo| <= 2^-20 |v|.
__device__ __forceinline__ void tf32_split(float v, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace msched
