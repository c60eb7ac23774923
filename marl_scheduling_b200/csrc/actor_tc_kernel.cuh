// actor_tc_kernel.cuh -- ActorCritic.act (src/PPOmodules.py:32-39,53-63) on the 5th-generation
// tensor cores: the three Linear layers of a 128-row tile are tcgen05.mma instructions with the
// accumulator in tensor memory, the Tanh / Softmax / Categorical epilogues run one row per thread.
//
//   rows     one (environment, unit) pair per row; a CTA of 128 threads owns 128 consecutive
//            environments of ONE unit (= one net), so the B operands are that net's weights
//   layer l  D[128 x N] = A[128 x K] * W^T, W in torch layout [N][K] = a K-major B operand
//   operands shared memory, canonical no-swizzle K-major layout: 8-row x 16-byte core matrices,
//            a K chunk (4 floats) of all rows is one panel; thread r writes its own row's 16 bytes
//            of a panel with one conflict-free 128-bit store
//   numerics the reference is fp32.  kind::tf32 keeps 11 significant bits, so every operand is
//            split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly)
//            and a product is three MMAs hi*hi + hi*lo + lo*hi ("3xTF32"): relative error about
//            2^-21 per product, the same order as the fp32 accumulation itself; observations are
//            small integers and split exactly
//   D        tensor memory, row i = lane i, column n; tcgen05.ld 32x32b gives thread i its row
//
// One elected thread issues the MMAs and commits them to an mbarrier; everybody waits on it,
// reads its accumulator row, applies bias + tanh and writes the next layer's A panels.
#pragma once
#include "msched_common.cuh"
#include "policy_kernels.cuh"
#include "tc_primitives.cuh"

namespace msched {

// how a warp reads the int16 observation rows of its 32 environments (see load_x_panels)
struct XLoadPlan {
    int WPR;   // aligned 32-bit words that can hold a row (either alignment)
    int wprp;  // WPR rounded up to a power of two (<= 32): lanes per row
    int rpi;   // rows per load instruction
};
__device__ __forceinline__ XLoadPlan x_load_plan(int nIn)
{
    XLoadPlan p;
    p.WPR = nIn / 2 + 1;
    p.wprp = 1;
    while (p.wprp < p.WPR && p.wprp < 32) p.wprp <<= 1;
    p.rpi = 32 / p.wprp;
    return p;
}

// Observation rows of a 128-environment tile -> layer-1 A panels (hi then lo, Kc1 chunks each) for
// unit `unit`.  Returns the core chooser's action of this thread's row in price-chooser mode (-1
// otherwise).  Called by all 128 threads of the CTA.
__device__ __forceinline__ int load_x_panels(const ActorArgs &a, unsigned char *aPan, int Kc1, int nIn, int tile,
                                             int unit, const XLoadPlan &pl)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int WPR = pl.WPR, wprp = pl.wprp, rpi = pl.rpi;
    const int env = tile * 128 + tid;
    const bool live = env < a.nEnvs;
    const int16_t *xr = a.x + (size_t)(live ? env : 0) * a.envStride + (size_t)unit * a.unitStride;
    const long long row = (long long)env * a.units + unit;
    int gsel = -1;
    if (a.gatherCore) {  // FreePriceOfferPPO.selectAction, src/PPOmodules.py:312-332 (quirk Q1)
        int16_t g4[4] = {0, 0, 0, 0};
        if (live) {
            gsel = a.gatherCore[row];
            const int c2 = 2 * a.nCores;
            if (gsel <= 0 || gsel > a.nCores) {
                g4[0] = g4[1] = g4[2] = g4[3] = (int16_t)-5;
            } else {
                g4[0] = xr[2 * gsel]; g4[1] = xr[2 * gsel + 1]; g4[2] = xr[c2]; g4[3] = xr[c2 + 1];
            }
            if (a.xUsed) *reinterpret_cast<short4 *>(a.xUsed + (size_t)row * 4) = make_short4(g4[0], g4[1], g4[2], g4[3]);
        }
        float v[4], hi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = (float)g4[q]; hi[q] = tf32_hi(v[q]); }
        *reinterpret_cast<float4 *>(aPan + tid * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(aPan + Kc1 * 2048 + tid * 16) =
            make_float4(v[0] - hi[0], v[1] - hi[1], v[2] - hi[2], v[3] - hi[3]);
        for (int c = 1; c < Kc1; ++c) {  // n_in == 4: the second chunk of the K = 8 step is zero
            *reinterpret_cast<float4 *>(aPan + c * 2048 + tid * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(aPan + (Kc1 + c) * 2048 + tid * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        // Warp-cooperative, coalesced: a row's int16 values are contiguous, so lane = (row of a
        // group of rpi rows, aligned 32-bit word of the row) reads whole sectors instead of
        // 32 lanes x 2 bytes from 32 different rows; the two values of a word are split and
        // written straight to their (row, k) place in the hi / lo panels
        const int lr = lane / wprp;
        for (int wb = 0; wb < WPR; wb += 32)  // more than one pass only for rows of 64+ values
        for (int r0 = 0; r0 < 32; r0 += 8 * rpi) {
            const int lw = wb + lane - lr * wprp;
            uint32_t wv[8];
            int k0v[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int rloc = warp * 32 + r0 + b * rpi + lr;
                const int e2 = tile * 128 + rloc;
                const int16_t *p = a.x + (size_t)(e2 < a.nEnvs ? e2 : 0) * a.envStride + (size_t)unit * a.unitStride;
                const int sh = (int)((reinterpret_cast<uintptr_t>(p) >> 1) & 1u);
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(p - sh);
                const int k0 = 2 * lw - sh;
                k0v[b] = k0;
                wv[b] = 0u;
                if (r0 + b * rpi + lr < 32 && b * rpi < 32 && e2 < a.nEnvs && lw < WPR && k0 < nIn) wv[b] = wp[lw];
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int rl = r0 + b * rpi + lr;
                if (rl < 32 && b * rpi < 32 && lw < WPR) {
                    const int rloc = warp * 32 + rl;
                    const long long row2 = (long long)(tile * 128 + rloc) * a.units + unit;
                    const bool rlive = tile * 128 + rloc < a.nEnvs;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int k = k0v[b] + hf;
                        if (k >= 0 && k < nIn) {
                            const int16_t e = (int16_t)(hf ? (wv[b] >> 16) : (wv[b] & 0xffffu));
                            const float v = (float)e, hi = tf32_hi(v);
                            const int off = rloc * 16 + (k & 3) * 4;
                            *reinterpret_cast<float *>(aPan + (k >> 2) * 2048 + off) = hi;
                            *reinterpret_cast<float *>(aPan + (Kc1 + (k >> 2)) * 2048 + off) = v - hi;
                            if (a.xUsed && rlive) a.xUsed[(size_t)row2 * nIn + k] = e;
                        }
                    }
                }
            }
        }
        // zero padding of this thread's own row (columns n_in .. 4*Kc1-1)
        for (int k = nIn; k < Kc1 * 4; ++k) {
            const int off = tid * 16 + (k & 3) * 4;
            *reinterpret_cast<float *>(aPan + (k >> 2) * 2048 + off) = 0.f;
            *reinterpret_cast<float *>(aPan + (Kc1 + (k >> 2)) * 2048 + off) = 0.f;
        }
    }
    return gsel;
}

// shared memory of one CTA: A panels (hi then lo), weights hi/lo per layer, biases
struct ActorTcSmem {
    int Kc1, Kc, aBytes, w1, w2, w3, bias, total;
    int Apad;
};
__host__ __device__ inline ActorTcSmem actor_tc_smem(int nIn, int H, int A)
{
    ActorTcSmem s;
    s.Kc1 = (nIn + 7) / 8 * 2;  // K chunks (4 floats) of layer 1, a multiple of 2 (UMMA K = 8)
    s.Kc = H / 4;
    s.Apad = (A + 15) / 16 * 16;
    const int kcMax = s.Kc1 > s.Kc ? s.Kc1 : s.Kc;
    s.aBytes = 2 * kcMax * 2048;
    s.w1 = s.aBytes;
    s.w2 = s.w1 + 2 * s.Kc1 * H * 16;
    s.w3 = s.w2 + 2 * s.Kc * H * 16;
    s.bias = s.w3 + 2 * s.Kc * s.Apad * 16;
    s.total = s.bias + (2 * H + s.Apad) * 4;
    return s;
}

template <int H, int AP>
__global__ void __launch_bounds__(128) actor_forward_tc(const ActorArgs a)
{
    extern __shared__ __align__(128) unsigned char smc[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int nIn = a.nIn, A = a.nActions;
    const ActorTcSmem L = actor_tc_smem(nIn, H, A);
    constexpr int Apad = AP;  // == L.Apad (checked by the launcher)
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + H * H + H + A * H + A;
    const float *w = a.weights + (size_t)net * pc;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t kCols = (H > 32 || AP > 32) ? 64u : 32u;  // accumulator columns (power of two >= 32)
    unsigned char *aPan = smc;
    float *bia = reinterpret_cast<float *>(smc + L.bias);

    unsigned long long *tl = (a.timeline && tid == 0) ? a.timeline + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (tl) { tl[0] = clock64(); tl[7] = smid(); }
    if (warp == 0) tmem_alloc(&tmemBase, kCols);
    if (tid == 32) mbar_init(&bar, 1);

    // ---- weights -> B operands (hi / lo), biases ----
    const float *w2 = w + H * nIn + H, *w3 = w2 + H * H + H;
    stage_weight(w, H, nIn, H, L.Kc1 * 4, smc + L.w1, smc + L.w1 + L.Kc1 * H * 16);
    stage_weight(w2, H, H, H, H, smc + L.w2, smc + L.w2 + L.Kc * H * 16);
    stage_weight(w3, A, H, Apad, H, smc + L.w3, smc + L.w3 + L.Kc * Apad * 16, kLog2e);  // base-2 logits
    for (int i = tid; i < H; i += 128) { bia[i] = w[H * nIn + i]; bia[H + i] = w2[H * H + i]; }
    for (int i = tid; i < Apad; i += 128) bia[2 * H + i] = i < A ? w3[A * H + i] * kLog2e : -INFINITY;  // padded logits

    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmemBase;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes
    const uint32_t aAddr = smem_u32(aPan);

    // issue one layer: D[128 x Npad] = A[128 x 4*Kc] * W^T with the three tf32 partial products
    auto issue_layer = [&](int Kc, int Npad, int wOff) {
        const uint32_t idesc = umma_idesc_tf32(128, Npad);
        const uint32_t bHi = smem_u32(smc + wOff), bLo = bHi + (uint32_t)(Kc * Npad * 16);
        const uint32_t aLo = aAddr + (uint32_t)(Kc * 2048);
        const uint32_t bLbo = (uint32_t)(Npad * 16);
        uint32_t acc = 0u;
        for (int ks = 0; ks < Kc / 2; ++ks) {
            const uint64_t ah = umma_smem_desc(aAddr + ks * 4096, 2048u, 128u);
            const uint64_t al = umma_smem_desc(aLo + ks * 4096, 2048u, 128u);
            const uint64_t bh = umma_smem_desc(bHi + ks * 2 * bLbo, bLbo, 128u);
            const uint64_t bl = umma_smem_desc(bLo + ks * 2 * bLbo, bLbo, 128u);
            umma_tf32(tbase, ah, bh, idesc, acc);
            umma_tf32(tbase, ah, bl, idesc, 1u);
            umma_tf32(tbase, al, bh, idesc, 1u);
            acc = 1u;
        }
        umma_commit(&bar);
    };
    // h -> hi / lo A panels of the next layer (its K = H)
    auto store_hidden = [&](const float (&h)[H]) {
#pragma unroll
        for (int c = 0; c < H / 4; ++c) {
            float hi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hi[q] = tf32_hi(h[4 * c + q]);
            *reinterpret_cast<float4 *>(aPan + c * 2048 + tid * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(aPan + (H / 4 + c) * 2048 + tid * 16) =
                make_float4(h[4 * c] - hi[0], h[4 * c + 1] - hi[1], h[4 * c + 2] - hi[2], h[4 * c + 3] - hi[3]);
        }
    };

    // ---- persistent loop over this unit's 128-row tiles ----
    const XLoadPlan xplan = x_load_plan(nIn);
    const int nTiles = (a.nEnvs + 127) / 128;
    uint32_t ph = 0u;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const bool stamp = tl && tile == (int)blockIdx.x;
        if (stamp) tl[1] = clock64();
        const int env = tile * 128 + tid;
        const bool live = env < a.nEnvs;
        const long long row = (long long)env * a.units + unit;
        const int gsel = load_x_panels(a, aPan, L.Kc1, nIn, tile, unit, xplan);
        fence_async_smem();
        tc_fence_before();
        if (stamp) tl[2] = clock64();
        __syncthreads();  // panels written; everybody has read the previous tile's accumulator
        tc_fence_after();

        // ---- layer 1 ----
        if (tid == 0) issue_layer(L.Kc1, H, L.w1);
        mbar_wait_bounded(&bar, ph);
        ph ^= 1u;
        tc_fence_after();
        float h[H];
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
            float v[16];
            tmem_ld16(trow + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bia + c * 16 + i);
                h[c * 16 + i] = fast_tanh(v[i] + b4.x); h[c * 16 + i + 1] = fast_tanh(v[i + 1] + b4.y);
                h[c * 16 + i + 2] = fast_tanh(v[i + 2] + b4.z); h[c * 16 + i + 3] = fast_tanh(v[i + 3] + b4.w);
            }
        }
        store_hidden(h);  // the layer-1 MMAs have completed: the panels are free
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (stamp) tl[3] = clock64();
        // ---- layer 2 ----
        if (tid == 0) issue_layer(L.Kc, H, L.w2);
        mbar_wait_bounded(&bar, ph);
        ph ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
            float v[16];
            tmem_ld16(trow + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bia + H + c * 16 + i);
                h[c * 16 + i] = fast_tanh(v[i] + b4.x); h[c * 16 + i + 1] = fast_tanh(v[i + 1] + b4.y);
                h[c * 16 + i + 2] = fast_tanh(v[i + 2] + b4.z); h[c * 16 + i + 3] = fast_tanh(v[i + 3] + b4.w);
            }
        }
        store_hidden(h);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (stamp) tl[4] = clock64();
        // ---- layer 3 (logits) ----
        if (tid == 0) issue_layer(L.Kc, Apad, L.w3);
        mbar_wait_bounded(&bar, ph);
        ph ^= 1u;
        tc_fence_after();
        float lg[AP];
#pragma unroll
        for (int c = 0; c < AP / 16; ++c) {
            float v[16];
            tmem_ld16(trow + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bia + 2 * H + c * 16 + i);
                lg[c * 16 + i] = v[i] + b4.x; lg[c * 16 + i + 1] = v[i + 1] + b4.y;
                lg[c * 16 + i + 2] = v[i + 2] + b4.z; lg[c * 16 + i + 3] = v[i + 3] + b4.w;
            }
        }
        if (stamp) tl[5] = clock64();
        // ---- Softmax, Categorical.sample (inverse CDF), Categorical.log_prob ----
        if (live) actor_epilogue(a, lg, A, row, env, unit, gsel);
        if (stamp) tl[6] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kCols);
}

}  // namespace msched
