// policy_step_tc_kernel.cuh -- msched_policy_step on the 5th-generation tensor cores.
//
// Same work and same contract as policy_step_kernel.cuh (every PPO unit of a rollout step in one launch:
// src/PPOmodules.py:32-39,53-63,114-125,312-332).  The three Linear layers of a 128-row tile are tcgen05.mma
// instructions; the threads only do what a matrix unit cannot: int16 -> fp16, Tanh, the hi/lo operand split,
// Softmax / Categorical.sample / log_prob.
//
//   CTA       serves ONE unit (an acceptor unit, or an offer unit = core chooser followed by the price chooser),
//             nets staged once as B operands in shared memory; SLOTS tile slots of 128 environments are in flight
//   warps     4 warps per slot (thread = row of the slot's tile = one environment); the slots of a CTA never meet.
//             There is no CTA barrier in the loop: a slot's 128 threads meet at the slot's NAMED barrier once their
//             rows of the next A operand are written, one elected thread of the slot issues the layer's MMAs and
//             commits them to the slot's `done` mbarrier, on which the slot's threads sleep.  While one slot waits
//             for its MMAs the other slots' warps run their epilogues
//   operands  BOTH the accumulator and the A operand live in TENSOR MEMORY (tcgen05.mma with A from tensor memory):
//             thread r owns lane r -- it reads its accumulator row with tcgen05.ld and writes its row of the next
//             layer's A operand with tcgen05.st, two fp16 per 32-bit column.  Nothing of the activation path goes
//             through shared memory, so the loop has no generic -> async proxy fence (the fence is a MEMBAR that
//             also waits for the thread's global loads and stores: with A panels in shared memory every layer of
//             every tile stalled on the tile's row loads / experience stores -- measured 59 us, and 54 us with the
//             MMAs taken out entirely), and an MMA no longer spends ~58 cycles reading a 4 KB A panel
//   numerics  fp16 PAIRS (kind::f16, fp32 accumulate): every fp32 value is split v = hi + lo with hi = v cut to 11
//             significant bits (exact in fp16) and lo = v - hi rounded to fp16: 22 bits, the products of the parts
//             are exact in the fp32 accumulator.  One MMA covers K = 16, i.e. a whole hidden layer
//   layer 1   the inputs are small integers (|x| <= 511: int16 -> fp16 exactly, by a mantissa trick on the integer
//             pipe), so A needs no split: D = X * W1hi^T + X * W1lo^T, two MMAs per 16 inputs
//   layers 2,3  h = hi + lo, D = Hhi*Whi^T + Hhi*Wlo^T + Hlo*Whi^T: three MMAs (about 2^-21 relative per product,
//             the order of the fp32 accumulation itself)
//   bias      one more MMA per layer: a constant A operand [1 1 0 ...] against a B chunk [b_hi b_lo 0 ...]
//             (accumulate = 0: it also initialises the accumulator), so the epilogues neither load nor add biases
//   scales    2*log2(e) folded into W1, b1, W2, b2 and log2(e) into W3, b3 (base-2 logits for the softmax)
//   tanh      1 - 2/(2^z' + 1); with a price chooser in the launch the reciprocals of FOUR values come from ONE rcp
//             (1/a = b*c*d / (a*b*c*d), z' clamped to 30 so that the product stays finite; tanh is 1.0f there
//             anyway): 20 instead of 32 transcendental-unit operations per row and layer (that unit does 16 lanes
//             per clock and SM); the other shapes take one rcp per value (fewer instructions), see tanh4_scaled
//   draws     one Philox call serves an environment pair: the two lanes of a pair compute the calls of two
//             consecutive tiles of the slot and swap words (tc_pair_draw)
//   loads     the observation rows of a warp's 32 environments are read warp-cooperatively (8 lanes per 32-byte
//             row: whole sectors) one tile AHEAD into registers, converted and transposed through a warp-private
//             staging tile to the row's owner; the same lanes write the experience-buffer copy of the row
//   price chooser  its four inputs are picked out of the owner's staged row (already fp16) by the sampled core
//   bound     instruction issue: ~15,500 thread instructions per environment (36 Tanh layers x ~165, 18 sampling rows x
//             ~115) at 60 % of the issue slots; tensor pipe 22 %.  The launch lasts as long as its slowest slot, so the
//             host splits the CTAs between the unit kinds on whole-tile counts (msched_rollout.cu); DESIGN.md 5.0 has
//             the measured history, including the variants that lost (work queue, two tiles per slot, issuer warp)
#pragma once
#include "policy_step_kernel.cuh"
#include "tc_primitives.cuh"

namespace msched {

// instruction descriptor: D fp32, A and B fp16, both K-major, M x N tile (kind::f16, K = 16 per instruction)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (lane = row, two fp16 per column)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// B operands of one 16-wide net: per layer hi chunks | lo chunks | 2 bias chunks (a chunk = 8 K values of the 16
// output rows = 256 bytes of fp16; the bias chunks hold b_hi, b_lo at K positions 0, 1)
template <int KC1>  // K chunks (8 halves) of layer 1; even (UMMA K = 16 for fp16)
struct TcNetImage {
    static constexpr int kL1 = 0, kL1Lo = KC1 * 256, kL1Bias = 2 * KC1 * 256;
    static constexpr int kL2 = kL1Bias + 512, kL2Lo = kL2 + 512, kL2Bias = kL2 + 1024;
    static constexpr int kL3 = kL2Bias + 512, kL3Lo = kL3 + 512, kL3Bias = kL3 + 1024;
    static constexpr int kBytes = kL3Bias + 512;
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
        // bias chunks: K position 0 = hi, 1 = lo, 2..15 = 0; logits beyond the net's A actions get -60000 (2^x = 0)
        for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
            const int n = i >> 4, k = i & 15;
            const float b1 = w[16 * nIn + n] * s2, b2 = w2[256 + n] * s2, b3 = n < A ? w3[A * 16 + n] * kLog2e : -60000.f;
            const int off = (k >> 3) * 256 + n * 16 + (k & 7) * 2;
            const float h1 = tf32_hi(b1), h2 = tf32_hi(b2), h3 = tf32_hi(b3);
            *reinterpret_cast<unsigned short *>(s + kL1Bias + off) = k == 0 ? f16_bits(h1) : (k == 1 ? f16_bits(b1 - h1) : (unsigned short)0);
            *reinterpret_cast<unsigned short *>(s + kL2Bias + off) = k == 0 ? f16_bits(h2) : (k == 1 ? f16_bits(b2 - h2) : (unsigned short)0);
            *reinterpret_cast<unsigned short *>(s + kL3Bias + off) = k == 0 ? f16_bits(h3) : (k == 1 ? f16_bits(b3 - h3) : (unsigned short)0);
        }
    }
};

// one layer of one slot: D[128 x 16] = ONES * BIAS + A[128 x 16*KS] * W^T with W = hi + lo; aLo == 0: exact A (two
// MMAs per K step of 16), else A = hi + lo as well (three).  A operands are tensor-memory columns (8 per K step).
// Issued by one thread, committed to the slot's `done` barrier
__device__ __forceinline__ void tc_issue_layer(uint32_t tmemD, uint32_t ones, uint32_t aHi, uint32_t aLo, uint32_t bHi, uint32_t bLo,
                                               uint32_t bBias, int KS, uint64_t *bar)
{
    constexpr uint32_t idesc = umma_idesc_f16(128, 16);
    umma_f16_ts(tmemD, ones, umma_smem_desc(bBias, 256u, 128u), idesc, 0u);
    for (int ks = 0; ks < KS; ++ks) {
        const uint64_t bh = umma_smem_desc(bHi + ks * 512, 256u, 128u);
        const uint64_t bl = umma_smem_desc(bLo + ks * 512, 256u, 128u);
        umma_f16_ts(tmemD, aHi + ks * 8, bh, idesc, 1u);
        umma_f16_ts(tmemD, aHi + ks * 8, bl, idesc, 1u);
        if (aLo) umma_f16_ts(tmemD, aLo + ks * 8, bh, idesc, 1u);
    }
    umma_commit(bar);
}

// the slot's 128 threads have written their rows of the next A operand into tensor memory and read the accumulator:
// they meet at the slot's named barrier, then one thread of the slot's first warp issues
__device__ __forceinline__ void tc_slot_issue(int slot, uint32_t tmemD, uint32_t ones, uint32_t aHi, uint32_t aLo, uint32_t bHi, uint32_t bLo,
                                              uint32_t bBias, int KS, uint64_t *done)
{
    tmem_st_wait();
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");
    if ((threadIdx.x & 127) < 32) {
        tc_fence_after();
        if ((threadIdx.x & 31) == 0) tc_issue_layer(tmemD, ones, aHi, aLo, bHi, bLo, bBias, KS, done);
        __syncwarp();
    }
}

// sleep on the slot's `done` barrier: one try_wait (the hardware suspends the thread up to the hint) almost always
// suffices; the retry loop is bounded and traps on a lost completion instead of hanging the device
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t addr, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(ns)
        : "memory");
    return ok;
}
__device__ __forceinline__ void tc_slot_wait(uint32_t doneAddr, uint32_t &k)
{
    const uint32_t parity = k & 1u;
    ++k;
    uint32_t ok = 0u;
#pragma unroll 1
    for (int it = 0; it < (1 << 16) && !ok; ++it) ok = mbar_try_wait_hint(doneAddr, parity, 20000u);
    if (!ok) __trap();
    tc_fence_after();
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// tanh of four pre-scaled values (z2 = 2*log2(e)*z): tanh z = 1 - 2 / (2^z2 + 1).
// MODE 0: one reciprocal per value -- four instructions and TWO special-function operations per value (2^z2 = inf
//   gives 1/inf = 0 -> +1, 2^z2 = 0 gives -1: no clamp).
// MODE 4: one reciprocal per four values -- 5.5 instructions and 1.25 special-function operations per value.  The
//   factors carry -1/2 each (folded into the "+1" as one FMA), so that the product's reciprocal comes out scaled by
//   the -2 of the formula: with e_i = -(2^z2_i + 1)/2, p = e0 e1, q = e2 e3, r = 1/(p q): 1 + e1 (q r) = 1 - 2/(2^z2_0 + 1).
// Which one is faster depends on how busy the special-function pipe already is: measured at 65,536 environments,
// the config-3 shape (price chooser: a third more softmax rows) 47.5 us with MODE 4 against 49.3 us with MODE 0, the
// config-2 shape 97.8 us against 93.5 us.
template <int MODE>
__device__ __forceinline__ void tanh4_scaled(float &x0, float &x1, float &x2, float &x3)
{
    if constexpr (MODE == 4) {
        const float e0 = fmaf(ex2_approx(fminf(x0, 30.f)), -0.5f, -0.5f), e1 = fmaf(ex2_approx(fminf(x1, 30.f)), -0.5f, -0.5f);
        const float e2 = fmaf(ex2_approx(fminf(x2, 30.f)), -0.5f, -0.5f), e3 = fmaf(ex2_approx(fminf(x3, 30.f)), -0.5f, -0.5f);
        const float p = e0 * e1, q = e2 * e3;
        const float r = rcp_approx(p * q);
        const float qr = q * r, pr = p * r;
        x0 = fmaf(e1, qr, 1.f);
        x1 = fmaf(e0, qr, 1.f);
        x2 = fmaf(e3, pr, 1.f);
        x3 = fmaf(e2, pr, 1.f);
    } else if constexpr (MODE == 2) {
        const float e0 = fmaf(ex2_approx(fminf(x0, 60.f)), -0.5f, -0.5f), e1 = fmaf(ex2_approx(fminf(x1, 60.f)), -0.5f, -0.5f);
        const float e2 = fmaf(ex2_approx(fminf(x2, 60.f)), -0.5f, -0.5f), e3 = fmaf(ex2_approx(fminf(x3, 60.f)), -0.5f, -0.5f);
        const float r = rcp_approx(e0 * e1), t = rcp_approx(e2 * e3);
        x0 = fmaf(e1, r, 1.f);
        x1 = fmaf(e0, r, 1.f);
        x2 = fmaf(e3, t, 1.f);
        x3 = fmaf(e2, t, 1.f);
    } else {
        x0 = fmaf(rcp_approx(ex2_approx(x0) + 1.f), -2.f, 1.f);
        x1 = fmaf(rcp_approx(ex2_approx(x1) + 1.f), -2.f, 1.f);
        x2 = fmaf(rcp_approx(ex2_approx(x2) + 1.f), -2.f, 1.f);
        x3 = fmaf(rcp_approx(ex2_approx(x3) + 1.f), -2.f, 1.f);
    }
}

// accumulator row (bias included) -> Tanh -> the thread's row of the next layer's hi / lo operands (8 + 8 columns)
template <int TANH>
__device__ __forceinline__ void tc_hidden_epilogue(uint32_t trow, uint32_t aHrow)
{
    float v[16];
    tmem_ld16(trow, v);
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float h0 = v[4 * c], h1 = v[4 * c + 1], h2 = v[4 * c + 2], h3 = v[4 * c + 3];
        tanh4_scaled<TANH>(h0, h1, h2, h3);
        const float i0 = tf32_hi(h0), i1 = tf32_hi(h1), i2 = tf32_hi(h2), i3 = tf32_hi(h3);
        hi[2 * c] = pack_f16x2(i0, i1); hi[2 * c + 1] = pack_f16x2(i2, i3);
        lo[2 * c] = pack_f16x2(h0 - i0, h1 - i1); lo[2 * c + 1] = pack_f16x2(h2 - i2, h3 - i3);
    }
    tmem_st8(aHrow, hi);
    tmem_st8(aHrow + 8, lo);
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

// logits of the thread's row (bias included, base 2) -> sample; AP = 8 or 16 columns are read
template <int AP, int AE>
__device__ __forceinline__ int tc_sample(uint32_t trow, int A, float u, float &logp, float *probsOut)
{
    float lg[AP];
    if constexpr (AP == 8) tmem_ld8(trow, lg);
    else tmem_ld16(trow, lg);
    return sample_row<AP, AE>(lg, A, u, logp, probsOut);
}

// The Philox draw of the thread's row.  One Philox call serves an environment PAIR (pair_draws: words 0, 1 = the two
// rows of the acceptor / core chooser, words 2, 3 = those of the price chooser), so the two lanes of a pair would
// compute the same ten rounds: instead, on an even iteration j of the slot's tile loop the even lane computes the
// call of THIS tile and the odd lane that of the slot's NEXT tile (tile + tileStep), and the two swap words with one
// shuffle per draw -- one Philox call per thread and TWO tiles, the draws themselves unchanged.  dr = the call the
// thread holds (kept across iterations).  Returns the row's word, the price chooser's word in vWord (WITH_V)
// SHARE false = every thread computes its own call (the 27-input acceptor shape has no registers to keep one across
// iterations: sharing measured 102.6 us against 97.5 us there, 48.9 against 51.4 us on the config-3 shape)
template <bool WITH_V, bool SHARE>
__device__ __forceinline__ uint32_t tc_pair_draw(const PolicyStepArgs &a, unsigned long long seed, int tile, int tileStep, int row,
                                                 int unit, int j, uint32_t (&dr)[4], uint32_t &vWord, unsigned long long stp)
{
    const int par = row & 1;  // = the lane's parity = the environment's parity (tiles start on even environments)
    if constexpr (!SHARE) {
        uint32_t r[4];
        pair_draws(a, seed, tile * 128 + row, unit, r, stp);
        if constexpr (WITH_V) vWord = par ? r[3] : r[2];
        return par ? r[1] : r[0];
    }
    if ((j & 1) == 0) pair_draws(a, seed, (tile + par * tileStep) * 128 + row, unit, dr, stp);
    const bool holder = par == (j & 1);  // this thread's call is the one of the current tile
    const uint32_t got = __shfl_xor_sync(0xffffffffu, par ? dr[0] : dr[1], 1);  // what the neighbour needs of mine
    const uint32_t u = holder ? (par ? dr[1] : dr[0]) : got;
    if constexpr (WITH_V) {
        const uint32_t gotV = __shfl_xor_sync(0xffffffffu, par ? dr[2] : dr[3], 1);
        vWord = holder ? (par ? dr[3] : dr[2]) : gotV;
    }
    return u;
}

// Observation rows of a warp's 32 environments.  LPR lanes per row read consecutive words: lane = (row slot rs, word w),
// iteration i covers rows i*RPI + rs.  fetch: global -> registers (one tile ahead); put: registers -> fp16 pairs (exact;
// the word's two values are K positions 2w, 2w+1 of the row) into the warp's staging tile [32 rows][SW words] and the
// experience buffer; the row's owner then moves its row into tensor memory.  The per-lane parts of the addresses
// are formed once per kernel, a tile adds its base and a row step per iteration
template <int KW, int SW>
struct TcRows {
    static constexpr int LPR = KW <= 2 ? 2 : KW <= 4 ? 4 : KW <= 8 ? 8 : KW <= 16 ? 16 : 32, RPI = 32 / LPR, NI = LPR;
    const uint32_t *src;   // obs + rs * strideW + offWords + w
    uint32_t *dst;         // xUsed + (rs * units + unit) * xuW + w, or null
    size_t srcEnv, dstEnv;  // words per environment
    int w, rs;
    __device__ __forceinline__ TcRows(const PolicyStepArgs &a, const PolicyGroupArgs &g, int unit, int offWords)
    {
        const int lane = threadIdx.x & 31;
        w = lane % LPR; rs = lane / LPR;
        srcEnv = (size_t)(a.obsStride >> 1);
        src = reinterpret_cast<const uint32_t *>(a.obs) + (size_t)rs * srcEnv + offWords + w;
        dstEnv = (size_t)g.units * (size_t)(g.xUsedStride >> 1);
        dst = g.xUsed ? reinterpret_cast<uint32_t *>(g.xUsed) + (size_t)rs * dstEnv + (size_t)unit * (g.xUsedStride >> 1) + w : nullptr;
    }
    __device__ __forceinline__ void fetch(int nEnvs, int env0, uint32_t (&v)[NI]) const
    {
        const uint32_t *p = src + (size_t)env0 * srcEnv;
        const size_t step = (size_t)RPI * srcEnv;
        if (w < KW) {
            if (env0 + 32 <= nEnvs) {
#pragma unroll
                for (int i = 0; i < NI; ++i) v[i] = __ldg(p + i * step);
            } else {
#pragma unroll
                for (int i = 0; i < NI; ++i) v[i] = (env0 + i * RPI + rs < nEnvs) ? __ldg(p + i * step) : 0u;
            }
        }
    }
    // stage: this warp's [32][SW] words
    __device__ __forceinline__ void put(int nEnvs, int env0, const uint32_t (&v)[NI], uint32_t *stage) const
    {
        if (w < KW) {
            uint32_t *sp = stage + rs * SW + w;
#pragma unroll
            for (int i = 0; i < NI; ++i) sp[i * RPI * SW] = halves_to_f16x2(v[i]);
            if (dst) {
                uint32_t *q = dst + (size_t)env0 * dstEnv;
                const size_t step = (size_t)RPI * dstEnv;
                if (env0 + 32 <= nEnvs) {
#pragma unroll
                    for (int i = 0; i < NI; ++i) q[i * step] = v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        if (env0 + i * RPI + rs < nEnvs) q[i * step] = v[i];
                }
            }
        }
    }
};

// the owner's staged row -> its lane of the layer-1 operand columns (SW words = SW / 8 K steps)
template <int SW>
__device__ __forceinline__ void tc_row_to_tmem(const uint32_t *stageRow, uint32_t aXrow)
{
#pragma unroll
    for (int q = 0; q < SW / 8; ++q) {
        const uint4 x0 = *reinterpret_cast<const uint4 *>(stageRow + 8 * q), x1 = *reinterpret_cast<const uint4 *>(stageRow + 8 * q + 4);
        const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        tmem_st8(aXrow + 8 * q, r);
    }
}

// KW_A / KW_O: words per acceptor / offer row; AP_*: logits columns read
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P, int SLOTS>
struct PolicyStepTcSmem {
    static constexpr int KS_A = (2 * KW_A + 15) / 16, KS_O = (2 * KW_O + 15) / 16;  // layer-1 K steps of 16
    static constexpr int kNetA = TcNetImage<2 * KS_A>::kBytes;
    static constexpr int kNetO = TcNetImage<2 * KS_O>::kBytes + (AP_P > 0 ? TcNetImage<2>::kBytes : 0);
    static constexpr int kNet = ((kNetA > kNetO ? kNetA : kNetO) + 127) & ~127;
    static constexpr int KSX = KS_A > KS_O ? KS_A : KS_O;
    static constexpr int SW = 8 * KSX;                      // staged words per row
    static constexpr int kStage = kNet;                     // [SLOTS * 4 warps][32 rows][SW] words
    static constexpr int kBytes = kStage + SLOTS * 128 * SW * 4;
    // tensor-memory columns of a slot: accumulator 16 | hidden hi 8 | hidden lo 8 | layer-1 operand 8 per K step
    static constexpr int kColsH = 16, kColsX = 32, kSlotCols = 32 + 8 * KSX;
    static constexpr int kColsOnes = SLOTS * kSlotCols;     // 8 columns [1 1 0 ...] shared by the slots
    static constexpr int kColsUsed = kColsOnes + 8;
    static constexpr uint32_t kTmemCols = kColsUsed <= 32 ? 32u : (kColsUsed <= 64 ? 64u : (kColsUsed <= 128 ? 128u : (kColsUsed <= 256 ? 256u : 512u)));
    static_assert(kColsUsed <= 512, "tensor memory has 512 columns");
};

// what one slot needs to run a net: its accumulator and operand columns, the CTA's constant operand
struct TcSlot {
    int slot;
    uint32_t tmemD, tOnes;  // lane 0 addresses (the issuing thread's view)
    uint32_t trow;          // the thread's lane of the slot's columns
    uint64_t *done;
    uint32_t doneAddr, k;
};

// layer 1 of a net (B image NI at shared address sNet): the operand is in the slot's layer-1 columns (ks1 K steps)
template <class NI, int COLS_X>
__device__ __forceinline__ void tc_issue_l1(TcSlot &t, uint32_t sNet, int ks1)
{
    tc_slot_issue(t.slot, t.tmemD, t.tOnes, t.tmemD + COLS_X, 0u, sNet + NI::kL1, sNet + NI::kL1Lo, sNet + NI::kL1Bias, ks1, t.done);
}

// the rest of the net for the thread's row, layer 1 being under way; returns the action
template <int AP, class NI, int TANH, int AE>
__device__ __forceinline__ int tc_run_net(TcSlot &t, uint32_t sNet, int A, float u, float &logp, float *probsOut)
{
    tc_slot_wait(t.doneAddr, t.k);
    tc_hidden_epilogue<TANH>(t.trow, t.trow + 16);
    tc_slot_issue(t.slot, t.tmemD, t.tOnes, t.tmemD + 16, t.tmemD + 24, sNet + NI::kL2, sNet + NI::kL2Lo, sNet + NI::kL2Bias, 1, t.done);
    tc_slot_wait(t.doneAddr, t.k);
    tc_hidden_epilogue<TANH>(t.trow, t.trow + 16);
    tc_slot_issue(t.slot, t.tmemD, t.tOnes, t.tmemD + 16, t.tmemD + 24, sNet + NI::kL3, sNet + NI::kL3Lo, sNet + NI::kL3Bias, 1, t.done);
    tc_slot_wait(t.doneAddr, t.k);
    return tc_sample<AP, AE>(t.trow, A, u, logp, probsOut);
}

// EXACT != 0: the kernel is built for one set of action counts (acceptor | core << 8 | price << 16) and its sampling
// epilogues skip the padding columns; 0: any counts up to the padded AP_*
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P, int SLOTS, int MINB, int EXACT = 0>
__global__ void __launch_bounds__(SLOTS * 128, MINB) policy_step_tc_kernel(const __grid_constant__ PolicyStepArgs a)
{
    using SM = PolicyStepTcSmem<KW_A, AP_A, KW_O, AP_O, AP_P, SLOTS>;
    using NA = TcNetImage<2 * SM::KS_A>;
    using NO = TcNetImage<2 * SM::KS_O>;
    using NP = TcNetImage<2>;
#ifndef MSCHED_TANH_PRICE_MODE
#define MSCHED_TANH_PRICE_MODE 4
#endif
    constexpr int AE_A = EXACT ? (EXACT & 0xff) : AP_A, AE_O = EXACT ? ((EXACT >> 8) & 0xff) : AP_O, AE_P = EXACT ? ((EXACT >> 16) & 0xff) : AP_P;
    constexpr int kTanh = AP_P > 0 ? MSCHED_TANH_PRICE_MODE : 0;  // tanh4_scaled: by the load of the special-function pipe
    constexpr int SW = SM::SW;
    extern __shared__ __align__(128) unsigned char smc[];
    __shared__ __align__(8) uint64_t barDone[SLOTS];
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
    // zeroed staging tiles: words beyond a row's KW stay zero for the whole kernel
    for (int i = tid; i < SLOTS * 128 * SW / 4; i += blockDim.x) reinterpret_cast<uint4 *>(smc + SM::kStage)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tmem_alloc(&tmemBase, SM::kTmemCols);
    if (tid == 32) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) mbar_init(&barDone[s], 1);
    }
    fence_async_smem();  // the B images are read by the MMAs (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int wq = warp & 3;
    const uint32_t laneSel = (uint32_t)(wq * 32) << 16;
    if (warp < 4) {  // the constant bias operand: fp16 1.0 (0x3c00) at K positions 0, 1 of every row
        const uint32_t ones[8] = {0x3c003c00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(tmemBase + SM::kColsOnes + laneSel, ones);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int myTiles = slice < nTiles ? (nTiles - slice + stride - 1) / stride : 0;  // tiles slice, slice+stride, ...: slot s takes every SLOTS-th

    const unsigned long long stepNow = policy_step_now(a);  // read once: the counter only moves between launches
    TcSlot t;
    t.slot = warp >> 2;
    const int row = tid & 127;
    const uint32_t sNet = smem_u32(smc);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smc + SM::kStage) + warp * 32 * SW;  // this warp's rows
    const uint32_t *stageRow = stage + lane * SW;
    t.tmemD = tmemBase + t.slot * SM::kSlotCols;
    t.tOnes = tmemBase + SM::kColsOnes;
    t.trow = t.tmemD + laneSel;
    t.done = &barDone[t.slot];
    t.doneAddr = smem_u32(t.done);
    t.k = 0u;
    if (isAcc) {
        const PolicyGroupArgs &g = a.acc;
        const TcRows<KW_A, SW> rl(a, g, unit, (g.xOffset - (g.xOffset & 1) + unit * g.xStride) >> 1);
        uint32_t rows[TcRows<KW_A, SW>::NI];
        uint32_t dr[4] = {0u, 0u, 0u, 0u}, dummyV;
        if (t.slot < myTiles) rl.fetch(a.nEnvs, (slice + t.slot * stride) * 128 + wq * 32, rows);
        for (int j = 0; j * SLOTS + t.slot < myTiles; ++j) {
            const int tile = slice + (j * SLOTS + t.slot) * stride;
            const int env = tile * 128 + row;
            const bool live = env < a.nEnvs;
            __syncwarp();
            rl.put(a.nEnvs, tile * 128 + wq * 32, rows, stage);
            __syncwarp();
            tc_row_to_tmem<8 * SM::KS_A>(stageRow, t.trow + SM::kColsX);
            tc_issue_l1<NA, SM::kColsX>(t, sNet, SM::KS_A);
            if ((j + 1) * SLOTS + t.slot < myTiles) rl.fetch(a.nEnvs, (tile + SLOTS * stride) * 128 + wq * 32, rows);
            float u;
            if (g.uOverride) {
                u = live ? g.uOverride[(size_t)env * g.units + unit] : 0.f;
            } else {
                u = u24(tc_pair_draw<false, (KW_A <= 8)>(a, g.seed, tile, SLOTS * stride, row, unit, j, dr, dummyV, stepNow));
            }
            float lp;
            const int act = tc_run_net<AP_A, NA, kTanh, AE_A>(t, sNet, g.nActions, u, lp,
                                                 (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
            if (live) emit_row(a, g, env, unit, act, lp, act);
        }
    } else {
        const PolicyGroupArgs &g = a.core, &gp = a.price;
        const TcRows<KW_O, SW> rl(a, g, unit, (g.xOffset + unit * g.xStride) >> 1);
        uint32_t rows[TcRows<KW_O, SW>::NI];
        uint32_t dr[4] = {0u, 0u, 0u, 0u};
        if (t.slot < myTiles) rl.fetch(a.nEnvs, (slice + t.slot * stride) * 128 + wq * 32, rows);
        for (int j = 0; j * SLOTS + t.slot < myTiles; ++j) {
            const int tile = slice + (j * SLOTS + t.slot) * stride;
            const int env = tile * 128 + row;
            const bool live = env < a.nEnvs;
            __syncwarp();  // (every lane has read its price inputs from the staging tile)
            rl.put(a.nEnvs, tile * 128 + wq * 32, rows, stage);
            __syncwarp();
            tc_row_to_tmem<8 * SM::KS_O>(stageRow, t.trow + SM::kColsX);
            tc_issue_l1<NO, SM::kColsX>(t, sNet, SM::KS_O);
            if ((j + 1) * SLOTS + t.slot < myTiles) rl.fetch(a.nEnvs, (tile + SLOTS * stride) * 128 + wq * 32, rows);
            float u, v = 0.f;
            if (g.uOverride) {
                u = live ? g.uOverride[(size_t)env * g.units + unit] : 0.f;
                if (AP_P > 0 && gp.uOverride) v = live ? gp.uOverride[(size_t)env * gp.units + unit] : 0.f;
            } else {
                uint32_t vw;
                u = u24(tc_pair_draw<(AP_P > 0), (KW_A <= 8)>(a, g.seed, tile, SLOTS * stride, row, unit, j, dr, vw, stepNow));
                v = u24(vw);
            }
            float lp;
            const int c = tc_run_net<AP_O, NO, kTanh, AE_O>(t, sNet, g.nActions, u, lp,
                                               (g.probs && live) ? g.probs + ((size_t)env * g.units + unit) * g.nActions : nullptr);
            if (live) emit_row(a, g, env, unit, c, lp, c);
            if constexpr (AP_P > 0) {
                // FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): [core prio, core rem, slot prio, slot rem] of
                // the chosen core = words c and nCores of the thread's own staged row (fp16 pairs); core action 0 feeds
                // the dummy [-5,-5,-5,-5] and reports price -5 (quirk Q1)
                const bool dummy = c <= 0 || c > a.nCores;
                const uint32_t wc = stageRow[dummy ? 0 : c], ws = stageRow[a.nCores];
                const uint32_t in0 = dummy ? 0xc500c500u : wc, in1 = dummy ? 0xc500c500u : ws;  // fp16 -5
                if (gp.xUsed && live)
                    *reinterpret_cast<short4 *>(gp.xUsed + ((size_t)env * gp.units + unit) * gp.xUsedStride) =
                        make_short4((short)f16lo_to_float(in0), (short)f16hi_to_float(in0), (short)f16lo_to_float(in1), (short)f16hi_to_float(in1));
                // (the layer-1 MMAs of the core chooser completed long ago)
                const uint32_t px[8] = {in0, in1, 0u, 0u, 0u, 0u, 0u, 0u};
                tmem_st8(t.trow + SM::kColsX, px);
                tc_issue_l1<NP, SM::kColsX>(t, sNet + NO::kBytes, 1);
                float lq;
                const int b = tc_run_net<AP_P, NP, kTanh, (AE_P > 0 ? AE_P : 1)>(t, sNet + NO::kBytes, gp.nActions, v, lq,
                                                   (gp.probs && live) ? gp.probs + ((size_t)env * gp.units + unit) * gp.nActions : nullptr);
                if (live) emit_row(a, gp, env, unit, b, lq, c == 0 ? -5 : b);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmemBase, SM::kTmemCols);
}

}  // namespace msched
