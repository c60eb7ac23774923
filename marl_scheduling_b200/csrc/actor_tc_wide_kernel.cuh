// actor_tc_wide_kernel.cuh -- ActorCritic.act for the AGGREGATED action heads
// (AggregatedAcceptorPPO / AggregatedOfferPPO / FullyAggregatedPPO, src/PPOmodules.py:177-232:
// (NL+1)^C, (C+1)^L or their product actions, 343 ... 28,561 in the reference domains).
//
// Layers 1 and 2 are those of actor_forward_tc.  The last Linear is a real GEMM
// [128 rows x H] x [H x A]: it is cut into tiles of 128 actions, each one tcgen05.mma group into one
// of two 128-column accumulator buffers in tensor memory, so the tensor cores work on tile t+1
// while the 128 threads (one row each) read tile t.  The [rows x A] logits are NEVER written to
// memory: softmax + Categorical sampling are two streaming passes over the tiles
//   pass 0  running maximum m and running sum s = sum exp(l - m)          (online softmax)
//   pass 1  the logits are recomputed; the first action whose running sum of exp(l - m) exceeds
//           u * s is the sample (inverse CDF in the un-normalised domain), log p = l - m - log s
// Weight tiles are staged like the A panels: thread n owns weight row n of the tile (its 4H bytes
// are contiguous in the torch layout), splits it into tf32 hi / lo and writes 16-byte pieces that
// are conflict-free across the warp.  The global loads of tile t+2 are issued before the epilogue
// of tile t and consumed after it.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "actor_tc_kernel.cuh"

namespace msched {

// actions per tile: 128 (= threads per CTA), 64 for the 64-wide nets whose weight tiles would
// not fit in shared memory next to the layer-1/2 operands otherwise
__host__ __device__ constexpr int wide_nt(int H) { return H >= 64 ? 64 : 128; }

struct ActorWideSmem {
    int Kc1, Kc, w1, w2, w3[2], bias, b3[2], total;
};
__host__ __device__ inline ActorWideSmem actor_wide_smem(int nIn, int H)
{
    ActorWideSmem s;
    s.Kc1 = (nIn + 7) / 8 * 2;
    s.Kc = H / 4;
    const int kcMax = s.Kc1 > s.Kc ? s.Kc1 : s.Kc;
    s.w1 = 2 * kcMax * 2048;
    s.w2 = s.w1 + 2 * s.Kc1 * H * 16;
    s.w3[0] = s.w2 + 2 * s.Kc * H * 16;
    const int NT = wide_nt(H);
    s.w3[1] = s.w3[0] + 2 * s.Kc * NT * 16;
    s.bias = s.w3[1] + 2 * s.Kc * NT * 16;
    s.b3[0] = s.bias + 2 * H * 4;
    s.b3[1] = s.b3[0] + NT * 4;
    s.total = s.b3[1] + NT * 4;
    return s;
}

template <int H>
__global__ void __launch_bounds__(128) actor_forward_tc_wide(const ActorArgs a)
{
    constexpr int NT = wide_nt(H), Kc = H / 4;
    extern __shared__ __align__(128) unsigned char smc[];
    __shared__ __align__(8) uint64_t bar12, barT[2];
    __shared__ uint32_t tmemBase;
    const int nIn = a.nIn, A = a.nActions;
    const ActorWideSmem L = actor_wide_smem(nIn, H);
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + H * H + H + A * H + A;
    const float *w = a.weights + (size_t)net * pc;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t kCols = 256u;  // two 128-column accumulator buffers
    unsigned char *aPan = smc;
    float *bia = reinterpret_cast<float *>(smc + L.bias);

    if (warp == 0) tmem_alloc(&tmemBase, kCols);
    if (tid == 32) { mbar_init(&bar12, 1); mbar_init(&barT[0], 1); mbar_init(&barT[1], 1); }

    const float *w2 = w + H * nIn + H, *w3 = w2 + H * H + H, *b3 = w3 + (size_t)A * H;
    stage_weight(w, H, nIn, H, L.Kc1 * 4, smc + L.w1, smc + L.w1 + L.Kc1 * H * 16);
    stage_weight(w2, H, H, H, H, smc + L.w2, smc + L.w2 + Kc * H * 16);
    for (int i = tid; i < H; i += 128) { bia[i] = w[H * nIn + i]; bia[H + i] = w2[H * H + i]; }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmemBase;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t aAddr = smem_u32(aPan);
    const bool w3vec = (reinterpret_cast<uintptr_t>(w3) & 15u) == 0;  // float4 loads of the weight rows

    auto issue_layer = [&](int kc, int Npad, uint32_t bHi, uint32_t dcol, uint64_t *bar) {
        const uint32_t idesc = umma_idesc_tf32(128, Npad);
        const uint32_t bLo = bHi + (uint32_t)(kc * Npad * 16);
        const uint32_t aLo = aAddr + (uint32_t)(kc * 2048);
        const uint32_t bLbo = (uint32_t)(Npad * 16);
        uint32_t acc = 0u;
        for (int ks = 0; ks < kc / 2; ++ks) {
            const uint64_t ah = umma_smem_desc(aAddr + ks * 4096, 2048u, 128u);
            const uint64_t al = umma_smem_desc(aLo + ks * 4096, 2048u, 128u);
            const uint64_t bh = umma_smem_desc(bHi + ks * 2 * bLbo, bLbo, 128u);
            const uint64_t bl = umma_smem_desc(bLo + ks * 2 * bLbo, bLbo, 128u);
            umma_tf32(tbase + dcol, ah, bh, idesc, acc);
            umma_tf32(tbase + dcol, ah, bl, idesc, 1u);
            umma_tf32(tbase + dcol, al, bh, idesc, 1u);
            acc = 1u;
        }
        umma_commit(bar);
    };
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
    // weight row (n0 + tid) of the last layer and its bias -> registers
    auto load_w3_row = [&](int n0, float4 (&wr)[Kc], float &br) {
        const int n = n0 + tid;
        if (tid >= NT) return;
        if (n < A) {
            const float *src = w3 + (size_t)n * H;
            if (w3vec) {
#pragma unroll
                for (int c = 0; c < Kc; ++c) wr[c] = *reinterpret_cast<const float4 *>(src + 4 * c);
            } else {
#pragma unroll
                for (int c = 0; c < Kc; ++c) wr[c] = make_float4(src[4 * c], src[4 * c + 1], src[4 * c + 2], src[4 * c + 3]);
            }
            // the last layer works in base 2: logits * log2(e), so that the softmax needs ex2 only
#pragma unroll
            for (int c = 0; c < Kc; ++c) { wr[c].x *= kLog2e; wr[c].y *= kLog2e; wr[c].z *= kLog2e; wr[c].w *= kLog2e; }
            br = b3[n] * kLog2e;
        } else {
#pragma unroll
            for (int c = 0; c < Kc; ++c) wr[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            br = -INFINITY;  // padded actions: logit -inf
        }
    };
    // registers -> B operand buffer `buf` (hi / lo panels of 128 rows x 16 bytes) + bias
    auto store_w3_row = [&](int buf, const float4 (&wr)[Kc], float br) {
        if (tid >= NT) return;
        unsigned char *dst = smc + L.w3[buf];
#pragma unroll
        for (int c = 0; c < Kc; ++c) {
            const float4 v = wr[c];
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *reinterpret_cast<float4 *>(dst + c * (NT * 16) + tid * 16) = hi;
            *reinterpret_cast<float4 *>(dst + (Kc + c) * (NT * 16) + tid * 16) =
                make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        }
        reinterpret_cast<float *>(smc + L.b3[buf])[tid] = br;
    };

    const XLoadPlan xplan = x_load_plan(nIn);
    const int nRowTiles = (a.nEnvs + 127) / 128;
    const int nT = (A + NT - 1) / NT;
    uint32_t ph12 = 0u, phT[2] = {0u, 0u};
    for (int tile = blockIdx.x; tile < nRowTiles; tile += gridDim.x) {
        const int env = tile * 128 + tid;
        const bool live = env < a.nEnvs;
        const long long row = (long long)env * a.units + unit;
        load_x_panels(a, aPan, L.Kc1, nIn, tile, unit, xplan);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // ---- layers 1 and 2 (accumulator: the first H columns of buffer 0) ----
        float h[H];
        if (tid == 0) issue_layer(L.Kc1, H, smem_u32(smc + L.w1), 0u, &bar12);
        mbar_wait_bounded(&bar12, ph12);
        ph12 ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
            float v[16];
            tmem_ld16(trow + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) h[c * 16 + i] = fast_tanh(v[i] + bia[c * 16 + i]);
        }
        store_hidden(h);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) issue_layer(Kc, H, smem_u32(smc + L.w2), 0u, &bar12);
        mbar_wait_bounded(&bar12, ph12);
        ph12 ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
            float v[16];
            tmem_ld16(trow + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) h[c * 16 + i] = fast_tanh(v[i] + bia[H + c * 16 + i]);
        }
        store_hidden(h);  // h2 stays in the A panels for every tile of the last layer

        // ---- the draw ----
        float u = 0.f;
        if (live) {
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
        }

        // ---- last layer: two streaming passes over the action tiles ----
        float m = -INFINITY, s = 0.f, thr = 0.f, cdf = 0.f, lact = 0.f, invs = 0.f;
        int act = -1;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) { thr = u * s; invs = 1.f / s; }
            // prologue: tiles 0 and 1 staged, their MMAs issued
            {
                float4 wr[Kc];
                float br;
                load_w3_row(0, wr, br);
                store_w3_row(0, wr, br);
                if (nT > 1) {
                    load_w3_row(NT, wr, br);
                    store_w3_row(1, wr, br);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();  // (also: h2 panels written, previous readers of both accumulators done)
            tc_fence_after();
            if (tid == 0) {
                issue_layer(Kc, NT, smem_u32(smc + L.w3[0]), 0u, &barT[0]);
                if (nT > 1) issue_layer(Kc, NT, smem_u32(smc + L.w3[1]), 128u, &barT[1]);
            }
            for (int t = 0; t < nT; ++t) {
                const int b = t & 1;
                float4 wr[Kc];
                float br = 0.f;
                const bool more = t + 2 < nT;
                if (more) load_w3_row((t + 2) * NT, wr, br);  // in flight during the epilogue below
                mbar_wait_bounded(&barT[b], phT[b]);
                phT[b] ^= 1u;
                tc_fence_after();
                const float *bt = reinterpret_cast<const float *>(smc + L.b3[b]);
#pragma unroll 1
                for (int c = 0; c < NT / 16; ++c) {
                    float v[16];
                    tmem_ld16(trow + (uint32_t)(b * 128 + c * 16), v);
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(bt + c * 16 + i);
                        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                    }
                    if (pass == 0) {
                        // online softmax in base 2: running max m, running sum s of 2^(l - m)
                        float cm = v[0];
#pragma unroll
                        for (int i = 1; i < 16; ++i) cm = fmaxf(cm, v[i]);
                        if (cm > m) { s *= ex2_approx(m - cm); m = cm; }  // (2^-inf = 0 on the first chunk)
                        float cs = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) cs += ex2_approx(v[i] - m);
                        s += cs;
                    } else {
                        const int col0 = t * NT + c * 16;
                        float e[16], cs = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) { e[i] = ex2_approx(v[i] - m); cs += e[i]; }
                        if (act < 0 && cdf + cs > thr) {  // the sample lies in this chunk (once per row)
                            float cc = cdf;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                cc += e[i];
                                const bool hit = act < 0 && cc > thr && col0 + i < A;
                                act = hit ? col0 + i : act;
                                lact = hit ? v[i] : lact;
                            }
                        }
                        cdf += cs;
                        if (a.probs && live) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (col0 + i < A) a.probs[(size_t)row * A + col0 + i] = e[i] * invs;
                        }
                    }
                }
                if (more) {
                    tc_fence_before();
                    __syncthreads();  // everybody has read accumulator b; MMA t has read weight buffer b
                    store_w3_row(b, wr, br);
                    fence_async_smem();
                    tc_fence_before();
                    __syncthreads();
                    tc_fence_after();
                    if (tid == 0) issue_layer(Kc, NT, smem_u32(smc + L.w3[b]), (uint32_t)(b * 128), &barT[b]);
                }
            }
            tc_fence_before();
            __syncthreads();  // all accumulator reads of this pass done before the next prologue's MMAs
            tc_fence_after();
        }
        if (live) {
            if (act < 0) {  // u * s rounded up to s: the last action; its logit is not at hand, use the
                act = A - 1;  // smallest representable probability bound below via the clamp
                lact = -INFINITY;
            }
            if (a.action) a.action[row] = act;
            if (a.actionRec) a.actionRec[(size_t)env * a.actionRecStride + unit] = (int16_t)act;
            if (a.logprob) {
                // Categorical.log_prob clamps the probability to [eps, 1-eps] before the log
                const float lo = -15.942385152878742f, hi = -1.1920929665620916e-07f;
                float lp = (lact - m) * kLn2 - logf(s);  // the logits are in base 2
                lp = fminf(fmaxf(lp, lo), hi);
                a.logprob[row] = lp;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kCols);
}

// 128-thread CTAs resident per SM: registers, shared memory and tensor-memory columns (the runtime's
// occupancy calculator reports 1 for kernels that allocate tensor memory, so it is derived here)
inline int resident_ctas(const void *fn, size_t dynSmem, int tmemCols)
{
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) return 1;
    const int regsPerThread = (fa.numRegs + 7) / 8 * 8;
    int byRegs = 65536 / (regsPerThread * 128);
    int bySmem = (int)((227 * 1024) / (dynSmem + fa.sharedSizeBytes + 1024));
    int byTmem = 512 / tmemCols;
    int r = byRegs < bySmem ? byRegs : bySmem;
    r = r < byTmem ? r : byTmem;
    return r < 1 ? 1 : r;
}

template <int H, int APS>
inline int launch_actor_simt(const ActorArgs &a, size_t smem, dim3 grid, cudaStream_t s)
{
    static int nSm = 0, perSm = 0;
    static size_t cachedSmem = ~(size_t)0;
    if (!nSm || cachedSmem != smem) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev);
        perSm = resident_ctas(reinterpret_cast<const void *>(actor_forward_simt<H, APS>), smem, 1);
        cachedSmem = smem;
    }
    int gx = (nSm * perSm) / (int)grid.y;  // persistent: one resident wave, weights staged once per CTA
    if (gx > (int)grid.x) gx = (int)grid.x;
    if (gx < 1) gx = 1;
    actor_forward_simt<H, APS><<<dim3(gx, grid.y), 128, smem, s>>>(a);
    return 0;
}

template <int H, int AP>
inline int launch_actor_shape(const ActorArgs &a, const MschedMlpGroup &g, dim3 grid, int impl, cudaStream_t s)
{
    if (impl == 0) {  // tensor cores (tcgen05, 3xTF32)
        const size_t smem = (size_t)actor_tc_smem(g.n_in, H, g.n_actions).total;
        if (smem > 200 * 1024) return -1;
        static size_t attrSmem = 0;
        if (smem > attrSmem) {
            if (cudaFuncSetAttribute(actor_forward_tc<H, AP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return -2;
            attrSmem = smem;
        }
        // persistent CTAs: exactly one resident wave (SMs x occupancy) shared by the units, each CTA
        // loops over its unit's tiles (weights staged and tensor memory allocated once per CTA)
        // (host queries cached per kernel instantiation and shared-memory size: a launch must stay
        // cheaper on the host than on the device)
        static int nSm = 0, perSm = 0;
        static size_t cachedSmem = 0;
        if (!nSm || cachedSmem != smem) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev);
            perSm = resident_ctas(reinterpret_cast<const void *>(actor_forward_tc<H, AP>), smem, (H > 32 || AP > 32) ? 64 : 32);
            if (getenv("MSCHED_DEBUG")) fprintf(stderr, "actor_forward_tc<%d,%d>: smem %zu resident CTAs/SM %d SMs %d\n", H, AP, smem, perSm, nSm);
            cachedSmem = smem;
        }
        int gx = (nSm * perSm) / (int)grid.y;
        if (gx > (int)grid.x) gx = (int)grid.x;
        if (gx < 1) gx = 1;
        actor_forward_tc<H, AP><<<dim3(gx, grid.y), 128, smem, s>>>(a);
        return 0;
    }
    if constexpr (H > 16) {
        return -1;  // the fp32 SIMT kernel is built for the 16-wide nets only (the wider nets run on the tensor cores)
    } else {
    const int Apad = (g.n_actions + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)g.n_in * H + H + (size_t)H * H + H + (size_t)H * Apad + Apad);
    // action buckets of the SIMT kernel: 8 (16-wide nets with up to 8 actions: half the softmax / sampling sweep
    // of the cfg3 acceptor and core chooser), 16, 64
    if (H == 16 && g.n_actions <= 8) return launch_actor_simt<H, 8>(a, smem, grid, s);
    return launch_actor_simt<H, (AP <= 16 ? 16 : 64)>(a, smem, grid, s);
    }
}

template <int H>
int launch_actor_h(const ActorArgs &a, const MschedMlpGroup &g, dim3 grid, int impl, cudaStream_t s)
{
    const int ap = (g.n_actions + 15) / 16 * 16;
    if (ap == 16) return launch_actor_shape<H, 16>(a, g, grid, impl, s);
    if (ap == 32) return launch_actor_shape<H, 32>(a, g, grid, impl, s);
    if (ap == 48) return launch_actor_shape<H, 48>(a, g, grid, impl, s);
    if (ap == 64) return launch_actor_shape<H, 64>(a, g, grid, impl, s);
    return -1;
}

// aggregated heads (more than 64 actions): tensor cores only, last layer tiled over the actions
template <int H>
int launch_actor_wide(const ActorArgs &a, const MschedMlpGroup &g, dim3 grid, cudaStream_t s)
{
    const size_t smem = (size_t)actor_wide_smem(g.n_in, H).total;
    if (smem > 220 * 1024) return -1;
    static size_t attrSmem = 0, cachedSmem = 0;
    static int nSm = 0, perSm = 0;
    if (smem > attrSmem) {
        if (cudaFuncSetAttribute(actor_forward_tc_wide<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -2;
        attrSmem = smem;
    }
    if (!nSm || cachedSmem != smem) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev);
        perSm = resident_ctas(reinterpret_cast<const void *>(actor_forward_tc_wide<H>), smem, 256);
        cachedSmem = smem;
    }
    int gx = (nSm * perSm) / (int)grid.y;
    if (gx > (int)grid.x) gx = (int)grid.x;
    if (gx < 1) gx = 1;
    actor_forward_tc_wide<H><<<dim3(gx, grid.y), 128, smem, s>>>(a);
    return 0;
}

// The kernel families are compiled in separate translation units (the build is as long as its slowest unit):
// msched_policy.cu holds the dispatcher and the 16-wide nets, msched_actor_h32.cu / _h64.cu the 32- / 64-wide nets,
// msched_actor_wide.cu the aggregated heads.  The dispatcher's unit sees the others as explicit instantiations
#ifdef MSCHED_ACTOR_DISPATCH_TU
extern template int launch_actor_h<32>(const ActorArgs &, const MschedMlpGroup &, dim3, int, cudaStream_t);
extern template int launch_actor_h<64>(const ActorArgs &, const MschedMlpGroup &, dim3, int, cudaStream_t);
extern template int launch_actor_wide<16>(const ActorArgs &, const MschedMlpGroup &, dim3, cudaStream_t);
extern template int launch_actor_wide<32>(const ActorArgs &, const MschedMlpGroup &, dim3, cudaStream_t);
extern template int launch_actor_wide<64>(const ActorArgs &, const MschedMlpGroup &, dim3, cudaStream_t);

inline ActorArgs make_actor_args(const MschedMlpGroup &g, const MschedActorIO &io)
{
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
    a.timeline = reinterpret_cast<unsigned long long *>(io.timeline);
    a.stepDev = reinterpret_cast<const unsigned long long *>(io.step_dev);
    return a;
}

inline int launch_actor_forward(const MschedMlpGroup &g, const MschedActorIO &io, int impl, cudaStream_t s)
{
    if (g.n_actions > 32767) return -1;  // actions are reported as int16 in the action record
    ActorArgs a = make_actor_args(g, io);
    dim3 grid((a.nEnvs + 127) / 128, io.units);
    if (g.n_actions > kActorMaxActions) {
        if (a.gatherCore) return -1;
        if (g.n_hidden == 16) return launch_actor_wide<16>(a, g, grid, s);
        if (g.n_hidden == 32) return launch_actor_wide<32>(a, g, grid, s);
        if (g.n_hidden == 64) return launch_actor_wide<64>(a, g, grid, s);
        return -1;
    }
    if (g.n_hidden == 16) return launch_actor_h<16>(a, g, grid, impl, s);
    if (g.n_hidden == 32) return launch_actor_h<32>(a, g, grid, impl, s);
    if (g.n_hidden == 64) return launch_actor_h<64>(a, g, grid, impl, s);
    return -1;
}
#endif  // MSCHED_ACTOR_DISPATCH_TU (the dispatcher exists in that unit only: an inline function that names every
        // instantiation makes the front end instantiate all of them in every unit that sees it)

}  // namespace msched
