// policy_step_tc_kernel.cuh -- msched_policy_step on the 5th-generation tensor cores, warp specialised.
//
// Same work and same contract as policy_step_kernel.cuh (every PPO unit of a rollout step in one launch:
// src/PPOmodules.py:32-39,53-63,114-125,312-332).  The three Linear layers of a 128-row tile are tcgen05.mma
// instructions with the accumulator in tensor memory; the threads only do what a matrix unit cannot:
// int16 -> float, Tanh, the hi/lo operand split, Softmax / Categorical.sample / log_prob.
//
//   CTA       serves ONE unit (an acceptor unit, or an offer unit = core chooser followed by the price chooser),
//             nets staged once as B operands; SLOTS tile slots of 128 environments are in flight at once
//   warps     4 epilogue warps per slot (thread = row of the slot's tile = one environment) + ONE issuer warp.
//             There is no CTA barrier in the loop: an epilogue warp that has written its rows of the next A operand
//             arrives on the slot's `ready` mbarrier (count 4); the issuer waits for it, issues the layer's MMAs and
//             commits them to the slot's `done` mbarrier, on which the slot's 128 threads sleep.  While one slot
//             waits for its MMAs the other slots' warps run their epilogues: the SM's issue slots and its
//             transcendental unit stay busy, and the MMA round trip (the whole cost of the first, serial version
//             of this kernel: 248 us) is hidden
//   operands  fp16 PAIRS in shared memory (kind::f16, fp32 accumulate), canonical no-swizzle K-major layout (8-row x
//             16-byte core matrices; a K chunk of 8 halves of all 128 rows is one 2 KB panel, thread r owns 16 bytes
//             of it: conflict-free 128-bit stores).  Every fp32 value is split v = hi + lo with hi = v cut to 11
//             significant bits (exact in fp16) and lo = v - hi rounded to fp16: 22 bits, the products of the parts
//             are exact in the fp32 accumulator.  One MMA covers K = 16, i.e. a whole hidden layer: the first
//             version of this kernel used kind::tf32 (K = 8, 32-bit operands) and needed 19 MMAs per acceptor
//             tile; ncu showed the tensor pipe 38 % busy at one CTA per SM -- each of these tiny MMAs holds it for
//             ~58 cycles, the time to read the 4 KB A panel from shared memory -- so the MMA COUNT was the bound
//   layer 1   the inputs are small integers (|x| <= 511: int16 -> fp16 exactly, by a mantissa trick on the integer
//             pipe), so A needs no split: D = X * W1hi^T + X * W1lo^T, two MMAs per 16 inputs
//   layers 2,3  h = hi + lo, D = Hhi*Whi^T + Hhi*Wlo^T + Hlo*Whi^T: three MMAs (about 2^-21 relative per product,
//             the order of the fp32 accumulation itself)
//   scales    2*log2(e) folded into W1, b1, W2, b2 and log2(e) into W3, b3 (base-2 logits for the softmax); the
//             biases are added by the epilogues (one more MMA per layer costs more than 16 FADDs)
//   tanh      1 - 2/(2^z' + 1); the reciprocals of FOUR values come from ONE rcp (1/a = b*c*d / (a*b*c*d), z'
//             clamped to 30 so that the product stays finite; tanh is 1.0f there anyway): 20 instead of 32
//             transcendental-unit operations per row and layer (that unit does 16 lanes per clock and SM)
//   loads     the observation rows of a warp's 32 environments are read warp-cooperatively (8 lanes per 32-byte
//             row: whole sectors) one tile AHEAD into registers, converted and written straight into the layer-1
//             A panels; the same lanes write the experience-buffer copy of the row
//   price chooser  its four inputs are picked out of the layer-1 A panels (already fp16) by the sampled core
#pragma once
#include "policy_step_kernel.cuh"
#include "tc_primitives.cuh"

namespace msched {

// instruction descriptor: D fp32, A and B fp16, both K-major, M x N tile (kind::f16, K = 16 per instruction)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// two floats -> one word of two fp16 (round to nearest), `even` in the low half (the lower K index)
__device__ __forceinline__ uint32_t pack_f16x2(float even, float odd)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
    return r;
}
__device__ __forceinline__ unsigned short f16_bits(float v) { return (unsigned short)(pack_f16x2(v, 0.f) & 0xffffu); }
// the two int16 halves of an observation word (|x| <= 511) as two fp16, exactly, without a conversion instruction:
// the low 10 bits with bit 9 flipped are x + 512; planted in the mantissa of fp16 1024.0 (ulp 1) they read 1536 + x
__device__ __forceinline__ uint32_t halves_to_f16x2(uint32_t w)
{
    const uint32_t m = (w & 0x03ff03ffu) ^ 0x66006600u;
    uint32_t r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0xe600e600u));  // - 1536
    return r;
}
__device__ __forceinline__ float f16lo_to_float(uint32_t w)
{
    float r;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l; }" : "=f"(r) : "r"(w));
    return r;
}
__device__ __forceinline__ float f16hi_to_float(uint32_t w)
{
    float r;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, h; }" : "=f"(r) : "r"(w));
    return r;
}

// B operands of one 16-wide net: per layer hi chunks | lo chunks (a chunk = 8 K values of the 16 output rows = 256
// bytes of fp16), then the scaled biases b1 | b2 | b3 as floats (b3 = -inf beyond the net's A actions)
template <int KC1>  // K chunks (8 halves) of layer 1; even (UMMA K = 16 for fp16)
struct TcNetImage {
    static constexpr int kL1 = 0, kL1Lo = KC1 * 256;
    static constexpr int kL2 = 2 * KC1 * 256, kL2Lo = kL2 + 512;
    static constexpr int kL3 = kL2 + 1024, kL3Lo = kL3 + 512;
    static constexpr int kBias = kL3 + 1024;
    static constexpr int kBytes = kBias + 48 * 4;
    __device__ static void put(unsigned char *hi, unsigned char *lo, int n, int k, float v)
    {
        const float h = tf32_hi(v);  // 13 low mantissa bits cleared: 11 significant bits, exact in fp16
        const int off = (k >> 3) * 256 + n * 16 + (k & 7) * 2;
        *reinterpret_cast<unsigned short *>(hi + off) = f16_bits(h);
        *reinterpret_cast<unsigned short *>(lo + off) = f16_bits(v - h);
    }
    // all threads; W1 row position = lead + input index (the int16 position inside the row words)
    __device__ static void stage(unsigned char *s, const float *__restrict__ w, int nIn, int lead, int A)
    {
        constexpr float s2 = 2.f * kLog2e;
        const float *w2 = w + 16 * nIn + 16, *w3 = w2 + 256 + 16;
        for (int i = threadIdx.x; i < 16 * KC1 * 8; i += blockDim.x) {
            const int n = i / (KC1 * 8), pos = i - n * (KC1 * 8), k = pos - lead;
            put(s + kL1, s + kL1Lo, n, pos, (k >= 0 && k < nIn) ? w[n * nIn + k] * s2 : 0.f);
        }
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            const int n = i >> 4, k = i & 15;
            put(s + kL2, s + kL2Lo, n, k, w2[n * 16 + k] * s2);
            put(s + kL3, s + kL3Lo, n, k, n < A ? w3[n * 16 + k] * kLog2e : 0.f);
        }
        float *b = reinterpret_cast<float *>(s + kBias);
        for (int i = threadIdx.x; i < 16; i += blockDim.x) {
            b[i] = w[16 * nIn + i] * s2;
            b[16 + i] = w2[256 + i] * s2;
            b[32 + i] = i < A ? w3[A * 16 + i] * kLog2e : -INFINITY;
        }
    }
};

// one layer of one slot: D[128 x 16] = A[128 x 8*KC] * W^T with W = hi + lo; aLo == 0: exact A (two MMAs per K step
// of 16), else A = hi + lo as well (three).  Issued by one thread, committed to the slot's `done` barrier
__device__ __forceinline__ void tc_issue_layer(uint32_t tmemD, uint32_t aHi, uint32_t aLo, uint32_t bHi, uint32_t bLo, int KC, uint64_t *bar)
{
    constexpr uint32_t idesc = umma_idesc_f16(128, 16);
    uint32_t acc = 0u;
    for (int ks = 0; ks < KC / 2; ++ks) {
        const uint64_t ah = umma_smem_desc(aHi + ks * 4096, 2048u, 128u);
        const uint64_t bh = umma_smem_desc(bHi + ks * 512, 256u, 128u);
        const uint64_t bl = umma_smem_desc(bLo + ks * 512, 256u, 128u);
        umma_f16(tmemD, ah, bh, idesc, acc);
        umma_f16(tmemD, ah, bl, idesc, 1u);
        if (aLo) umma_f16(tmemD, umma_smem_desc(aLo + ks * 4096, 2048u, 128u), bh, idesc, 1u);
        acc = 1u;
    }
    umma_commit(bar);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// an epilogue warp hands its rows of the slot to the issuer: operand stores visible to the async proxy,
// accumulator reads done, then one arrival per warp
__device__ __forceinline__ void tc_slot_arrive(uint64_t *ready)
{
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(ready);
}

__device__ __forceinline__ void tc_slot_wait(uint64_t *done, uint32_t &k)
{
    mbar_wait_bounded(done, k & 1u);
    ++k;
    tc_fence_after();
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// tanh of four pre-scaled values (z2 = 2*log2(e)*z) with one reciprocal
__device__ __forceinline__ void tanh4_scaled(float &x0, float &x1, float &x2, float &x3)
{
    const float e0 = ex2_approx(fminf(x0, 30.f)) + 1.f, e1 = ex2_approx(fminf(x1, 30.f)) + 1.f;
    const float e2 = ex2_approx(fminf(x2, 30.f)) + 1.f, e3 = ex2_approx(fminf(x3, 30.f)) + 1.f;
    const float p = e0 * e1, q = e2 * e3;
    const float r = rcp_approx(p * q);
    const float qr = q * r, pr = p * r;
    x0 = fmaf(-2.f, e1 * qr, 1.f);
    x1 = fmaf(-2.f, e0 * qr, 1.f);
    x2 = fmaf(-2.f, e3 * pr, 1.f);
    x3 = fmaf(-2.f, e2 * pr, 1.f);
}

// accumulator row + bias -> Tanh -> hi / lo A panels of the next layer (K = 16 halves: 2 + 2 panels of 2 KB)
__device__ __forceinline__ void tc_hidden_epilogue(uint32_t trow, const float *__restrict__ bias, unsigned char *aH, int row)
{
    float v[16];
    tmem_ld16(trow, v);
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + 4 * c);
        float h0 = v[4 * c] + b4.x, h1 = v[4 * c + 1] + b4.y, h2 = v[4 * c + 2] + b4.z, h3 = v[4 * c + 3] + b4.w;
        tanh4_scaled(h0, h1, h2, h3);
        const float i0 = tf32_hi(h0), i1 = tf32_hi(h1), i2 = tf32_hi(h2), i3 = tf32_hi(h3);
        hi[2 * c] = pack_f16x2(i0, i1); hi[2 * c + 1] = pack_f16x2(i2, i3);
        lo[2 * c] = pack_f16x2(h0 - i0, h1 - i1); lo[2 * c + 1] = pack_f16x2(h2 - i2, h3 - i3);
    }
    *reinterpret_cast<uint4 *>(aH + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(aH + 2048 + row * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4 *>(aH + 4096 + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4 *>(aH + 6144 + row * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// logits of the thread's row (base 2) -> sample; AP = 8 or 16 columns are read
template <int AP>
__device__ __forceinline__ int tc_sample(uint32_t trow, const float *__restrict__ b3, int A, float u, float &logp, float *probsOut)
{
    float lg[AP];
    if constexpr (AP == 8) tmem_ld8(trow, lg);
    else tmem_ld16(trow, lg);
#pragma unroll
    for (int o = 0; o < AP; ++o) lg[o] += b3[o];
    return sample_row<AP>(lg, A, u, logp, probsOut);
}

// Observation rows of a warp's 32 environments.  LPR lanes per row read consecutive words: lane = (row slot rs, word w),
// iteration i covers rows i*RPI + rs.  fetch: global -> registers (one tile ahead); put: registers -> layer-1 A panels
// (fp16, exact; the word's two values land at K positions 2w, 2w+1 of the row) and the experience buffer
template <int KW>
struct TcRows {
    static constexpr int LPR = KW <= 2 ? 2 : KW <= 4 ? 4 : KW <= 8 ? 8 : KW <= 16 ? 16 : 32, RPI = 32 / LPR, NI = LPR;
    __device__ static __forceinline__ void fetch(const PolicyStepArgs &a, int env0, int offWords, uint32_t (&v)[NI])
    {
        const int lane = threadIdx.x & 31, w = lane % LPR, rs = lane / LPR;
        const uint32_t *ob = reinterpret_cast<const uint32_t *>(a.obs);
        const long long strideW = a.obsStride >> 1;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int env = env0 + i * RPI + rs;
            v[i] = (w < KW && env < a.nEnvs) ? __ldg(ob + (size_t)env * strideW + offWords + w) : 0u;
        }
    }
    __device__ static __forceinline__ void put(const PolicyStepArgs &a, const PolicyGroupArgs &g, int env0, int unit, const uint32_t (&v)[NI],
                                               unsigned char *aX, int row0)
    {
        const int lane = threadIdx.x & 31, w = lane % LPR, rs = lane / LPR;
        uint32_t *xu = reinterpret_cast<uint32_t *>(g.xUsed);
        const int xuW = g.xUsedStride >> 1;
        if (w < KW) {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int r = i * RPI + rs, env = env0 + r;
                *reinterpret_cast<uint32_t *>(aX + (w >> 2) * 2048 + (row0 + r) * 16 + (w & 3) * 4) = halves_to_f16x2(v[i]);
                if (xu && env < a.nEnvs) xu[((size_t)env * g.units + unit) * xuW + w] = v[i];
            }
        }
    }
};

// KW_A / KW_O: words per acceptor / offer row; AP_*: logits columns read
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P, int SLOTS>
struct PolicyStepTcSmem {
    static constexpr int KC_A = ((2 * KW_A + 15) / 16) * 2, KC_O = ((2 * KW_O + 15) / 16) * 2;  // layer-1 K chunks of 8, even
    static constexpr int kNetA = TcNetImage<KC_A>::kBytes;
    static constexpr int kNetO = TcNetImage<KC_O>::kBytes + (AP_P > 0 ? TcNetImage<2>::kBytes : 0);
    static constexpr int kNet = ((kNetA > kNetO ? kNetA : kNetO) + 127) & ~127;
    static constexpr int KCX = KC_A > KC_O ? KC_A : KC_O;
    static constexpr int kSlot0 = kNet;
    static constexpr int kAH = KCX * 2048;                 // inside a slot: layer-1 panels | hidden hi (2) | hidden lo (2)
    static constexpr int kSlotBytes = kAH + 4 * 2048;
    static constexpr int kBytes = kSlot0 + SLOTS * kSlotBytes;
    static constexpr uint32_t kTmemCols = SLOTS * 16 <= 32 ? 32u : (SLOTS * 16 <= 64 ? 64u : 128u);
};

// the three layers of one net for the thread's row, given that the layer-1 operand has been handed over; returns the action
template <int AP>
__device__ __forceinline__ int tc_run_net(uint32_t trow, const float *__restrict__ bias, unsigned char *aH, int row, uint64_t *ready,
                                          uint64_t *done, uint32_t &k, int A, float u, float &logp, float *probsOut)
{
    tc_slot_wait(done, k);
    tc_hidden_epilogue(trow, bias, aH, row);
    tc_slot_arrive(ready);
    tc_slot_wait(done, k);
    tc_hidden_epilogue(trow, bias + 16, aH, row);
    tc_slot_arrive(ready);
    tc_slot_wait(done, k);
    return tc_sample<AP>(trow, bias + 32, A, u, logp, probsOut);
}

template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P, int SLOTS, int MINB>
__global__ void __launch_bounds__(SLOTS * 128 + 32, MINB) policy_step_tc_kernel(const __grid_constant__ PolicyStepArgs a)
{
    using SM = PolicyStepTcSmem<KW_A, AP_A, KW_O, AP_O, AP_P, SLOTS>;
    using NA = TcNetImage<SM::KC_A>;
    using NO = TcNetImage<SM::KC_O>;
    using NP = TcNetImage<2>;
    extern __shared__ __align__(128) unsigned char smc[];
    __shared__ __align__(8) uint64_t barReady[SLOTS], barDone[SLOTS];
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kIssuer = SLOTS * 4;
    const int nAccCtas = a.acc.units * a.ctasPerAccUnit;
    const bool isAcc = (int)blockIdx.x < nAccCtas;
    const int nTiles = (a.nEnvs + 127) / 128;

    int unit, slice, stride;
    if (isAcc) {
        unit = blockIdx.x / a.ctasPerAccUnit; slice = blockIdx.x - unit * a.ctasPerAccUnit; stride = a.ctasPerAccUnit;
        const PolicyGroupArgs &g = a.acc;
        const int net = (unit / g.unitDiv) % g.nNets;
        const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
        NA::stage(smc, g.weights + (size_t)net * pc, g.nIn, g.xOffset & 1, g.nActions);
    } else {
        const int id = blockIdx.x - nAccCtas;
        unit = id / a.ctasPerOffUnit; slice = id - unit * a.ctasPerOffUnit; stride = a.ctasPerOffUnit;
        const PolicyGroupArgs &g = a.core, &gp = a.price;
        const int net = (unit / g.unitDiv) % g.nNets;
        const int pc = 16 * g.nIn + 16 + 256 + 16 + 16 * g.nActions + g.nActions;
        NO::stage(smc, g.weights + (size_t)net * pc, g.nIn, 0, g.nActions);
        if constexpr (AP_P > 0) {
            const int netp = (unit / gp.unitDiv) % gp.nNets;
            const int pcp = 16 * 4 + 16 + 256 + 16 + 16 * gp.nActions + gp.nActions;
            NP::stage(smc + NO::kBytes, gp.weights + (size_t)netp * pcp, 4, 0, gp.nActions);
        }
    }
    // zeroed A panels: K positions beyond a row's words stay zero for the whole kernel
    for (int i = tid; i < SLOTS * SM::kSlotBytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smc + SM::kSlot0)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == kIssuer) tmem_alloc(&tmemBase, SM::kTmemCols);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) { mbar_init(&barReady[s], 4); mbar_init(&barDone[s], 1); }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmemBase;
    const int myTiles = slice < nTiles ? (nTiles - slice + stride - 1) / stride : 0;  // tiles slice, slice+stride, ...: slot s takes every SLOTS-th
    const int lpt = (!isAcc && AP_P > 0) ? 6 : 3;                                     // layers per tile

    if (warp == kIssuer) {
        const uint32_t sNet = smem_u32(smc), sSlot0 = sNet + SM::kSlot0;
        const int maxSteps = ((myTiles + SLOTS - 1) / SLOTS) * lpt;
        int li = 0, j = 0;
        for (int it = 0; it < maxSteps; ++it) {
            // operands of layer li (0..2 the unit's first net, 3..5 the price chooser)
            uint32_t aOff, aLo, b, bLo;
            int kc;
            if (li == 0 || li == 3) {
                aOff = 0u; aLo = 0u;
                if (isAcc) { b = NA::kL1; bLo = NA::kL1Lo; kc = SM::KC_A; }
                else if (li == 0) { b = NO::kL1; bLo = NO::kL1Lo; kc = SM::KC_O; }
                else { b = NO::kBytes + NP::kL1; bLo = NO::kBytes + NP::kL1Lo; kc = 2; }
            } else {
                aOff = SM::kAH; aLo = SM::kAH + 2 * 2048;
                const bool second = (li == 1 || li == 4);
                if (isAcc) { b = second ? NA::kL2 : NA::kL3; bLo = second ? NA::kL2Lo : NA::kL3Lo; }
                else if (li < 3) { b = second ? NO::kL2 : NO::kL3; bLo = second ? NO::kL2Lo : NO::kL3Lo; }
                else { b = NO::kBytes + (second ? NP::kL2 : NP::kL3); bLo = NO::kBytes + (second ? NP::kL2Lo : NP::kL3Lo); }
                kc = 2;
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                if (j * SLOTS + s < myTiles) {
                    mbar_wait_bounded(&barReady[s], (uint32_t)it & 1u);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sl = sSlot0 + s * SM::kSlotBytes;
                        tc_issue_layer(tbase + s * 16, sl + aOff, aLo ? sl + aLo : 0u, sNet + b, sNet + bLo, kc, &barDone[s]);
                    }
                    __syncwarp();
                }
            }
            if (++li == lpt) { li = 0; ++j; }
        }
    } else {
        const int slot = warp >> 2, wq = warp & 3, row = tid & 127;
        unsigned char *aX = smc + SM::kSlot0 + slot * SM::kSlotBytes, *aH = aX + SM::kAH;
        const uint32_t trow = tbase + slot * 16 + ((uint32_t)(wq * 32) << 16);
        uint64_t *ready = &barReady[slot], *done = &barDone[slot];
        uint32_t k = 0u;
        if (isAcc) {
            using RL = TcRows<KW_A>;
            const PolicyGroupArgs &g = a.acc;
            const float *bias = reinterpret_cast<const float *>(smc + NA::kBias);
            const int offW = (g.xOffset - (g.xOffset & 1) + unit * g.xStride) >> 1;
            uint32_t rows[RL::NI];
            if (slot < myTiles) RL::fetch(a, (slice + slot * stride) * 128 + wq * 32, offW, rows);
            for (int j = 0; j * SLOTS + slot < myTiles; ++j) {
                const int tile = slice + (j * SLOTS + slot) * stride;
                const int env = tile * 128 + row;
                const bool live = env < a.nEnvs;
                RL::put(a, g, tile * 128 + wq * 32, unit, rows, aX, wq * 32);
                tc_slot_arrive(ready);
                if ((j + 1) * SLOTS + slot < myTiles) RL::fetch(a, (tile + SLOTS * stride) * 128 + wq * 32, offW, rows);
                float u;
                if (g.uOverride) {
                    u = live ? g.uOverride[(size_t)env * g.units + unit] : 0.f;
                } else {
                    uint32_t r[4];
                    pair_draws(a, g.seed, env, unit, r);
                    u = u24((env & 1) ? r[1] : r[0]);
                }
                float lp;
                const int act = tc_run_net<AP_A>(trow, bias, aH, row, ready, done, k, g.nActions, u, lp,
                                                 (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
                if (live) emit_row(a, g, env, unit, act, lp, act);
            }
        } else {
            using RL = TcRows<KW_O>;
            const PolicyGroupArgs &g = a.core, &gp = a.price;
            const float *bias = reinterpret_cast<const float *>(smc + NO::kBias);
            const int offW = (g.xOffset + unit * g.xStride) >> 1;
            uint32_t rows[RL::NI];
            if (slot < myTiles) RL::fetch(a, (slice + slot * stride) * 128 + wq * 32, offW, rows);
            for (int j = 0; j * SLOTS + slot < myTiles; ++j) {
                const int tile = slice + (j * SLOTS + slot) * stride;
                const int env = tile * 128 + row;
                const bool live = env < a.nEnvs;
                RL::put(a, g, tile * 128 + wq * 32, unit, rows, aX, wq * 32);
                tc_slot_arrive(ready);
                if ((j + 1) * SLOTS + slot < myTiles) RL::fetch(a, (tile + SLOTS * stride) * 128 + wq * 32, offW, rows);
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
                float lp;
                const int c = tc_run_net<AP_O>(trow, bias, aH, row, ready, done, k, g.nActions, u, lp,
                                               (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
                if (live) emit_row(a, g, env, unit, c, lp, c);
                if constexpr (AP_P > 0) {
                    // FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): [core prio, core rem, slot prio, slot rem] of
                    // the chosen core = row words c and nCores of the thread's own layer-1 row (fp16 pairs); core action 0
                    // feeds the dummy [-5,-5,-5,-5] and reports price -5 (quirk Q1)
                    const float *biasP = reinterpret_cast<const float *>(smc + NO::kBytes + NP::kBias);
                    const bool dummy = c <= 0 || c > a.nCores;
                    const int cc = dummy ? 0 : c;
                    const uint32_t wc = *reinterpret_cast<const uint32_t *>(aX + (cc >> 2) * 2048 + row * 16 + (cc & 3) * 4);
                    const uint32_t ws = *reinterpret_cast<const uint32_t *>(aX + (a.nCores >> 2) * 2048 + row * 16 + (a.nCores & 3) * 4);
                    const uint32_t in0 = dummy ? 0xc500c500u : wc, in1 = dummy ? 0xc500c500u : ws;  // fp16 -5
                    if (gp.xUsed && live)
                        *reinterpret_cast<short4 *>(gp.xUsed + ((size_t)env * gp.units + unit) * gp.xUsedStride) =
                            make_short4((short)f16lo_to_float(in0), (short)f16hi_to_float(in0), (short)f16lo_to_float(in1), (short)f16hi_to_float(in1));
                    // (the layer-1 MMAs of the core chooser completed long ago; a thread rewrites only its OWN 16-byte slots)
                    *reinterpret_cast<uint4 *>(aX + row * 16) = make_uint4(in0, in1, 0u, 0u);
                    *reinterpret_cast<uint4 *>(aX + 2048 + row * 16) = make_uint4(0u, 0u, 0u, 0u);
                    tc_slot_arrive(ready);
                    float lq;
                    const int b = tc_run_net<AP_P>(trow, biasP, aH, row, ready, done, k, gp.nActions, v, lq,
                                                   (gp.probs && live) ? gp.probs + ((size_t)env * gp.units + unit) * gp.nActions : nullptr);
                    if (live) emit_row(a, gp, env, unit, b, lq, c == 0 ? -5 : b);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kIssuer) tmem_dealloc(tbase, SM::kTmemCols);
}

}  // namespace msched
