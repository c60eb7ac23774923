// ppo_update_kernel.cuh -- one epoch's gradient of PPO.update for a group of identically shaped
// ActorCritic nets, forward AND backward in one kernel.
//
//   src/PPOmodules.py:65-72     ActorCritic.evaluate: log-prob of the stored action, entropy, V(s)
//   src/PPOmodules.py:139-174   ratios, clipped surrogate, 0.5*MSE, -0.01*entropy, loss.mean().backward()
//   src/PPOmodules.py:100-105   torch.optim.Adam with one learning rate per head (adam_kernel)
//
// The nets are tiny (in <= 64, 16 hidden neurons, <= 16 actions) and the batch is huge (T*B samples per
// unit), so one THREAD owns one sample: it evaluates critic and actor in fp32 registers against weights
// staged in shared memory ([in][out] order: the same image serves x·W^T forwards and dy·W backwards),
// forms dL/dz of every layer and hands the per-sample outer products dW = sum_r dy[r] (x) act[r] to the
// tensor cores: the warp parks dy and act of its 32 samples in shared memory (element-major, stride 36:
// conflict-free for the per-thread stores and the fragment loads alike) and contracts them over the
// SAMPLE dimension with mma.sync.m16n8k8 TF32, every operand split hi+lo (3xTF32, fp32 accuracy),
// accumulators resident in registers across the whole persistent loop.  The reduction over samples that
// dominates a SIMT backward therefore costs 24 MMAs per 16x16 weight matrix and 32 samples.  Partial sums
// leave per CTA, and ppo_reduce_kernel adds them in a fixed order (bit-reproducible gradients, no atomics).
//
// Gradient per sample (M = samples of the net, applied in the reduce kernel):
//   ratio = exp(logp_a - logp_old), adv = G - V(s) (detached)
//   d = adv if ratio*adv < clamp(ratio)*adv or 1-eps <= ratio <= 1+eps, else 0   (torch.min / clamp subgradients)
//   dL/dz_j = -d*ratio*([j==a] - p_j) + c_ent*p_j*(log p_j + H)                  (actor logits)
//   dL/dV   = 2*c_val*(V - G)                                                    (critic, MseLoss is a batch mean)
#pragma once
#include "msched_common.cuh"
#include "policy_common.cuh"

#ifndef MSCHED_PPO_UNROLL
#define MSCHED_PPO_UNROLL 2  // unroll factor of the per-sample runtime sweeps (code size against loop overhead)
#endif
constexpr int kPpoUnroll = MSCHED_PPO_UNROLL;

namespace msched {

struct PpoArgs {
    const float *actorW, *criticW;  // [n_nets][pc]
    const int16_t *x;
    long long xTbStride, xUnitStride;  // int16 elements
    const int32_t *action;             // [TB][U]
    const float *logpOld, *ret;        // [TB][U]
    const int32_t *netIds;             // [nSel]
    const int32_t *unitIds;            // [nSel][m]
    long long nTb;
    int U, nSel, m, nIn, A;
    float epsClip, entCoef, valCoef;
    float *partial;  // [nSel][gridDim.x][P], P = pcA + pcC + 4
};

constexpr int kPpoH = 16;
constexpr int kPpoStride = 36;  // samples per warp tile (32) + 4: (4*e + r) mod 32 distinct for fragment loads

__host__ __device__ inline int ppo_param_count(int nIn, int A) { return kPpoH * nIn + kPpoH + kPpoH * kPpoH + kPpoH + A * kPpoH + A; }

// one net's shared-memory image, weights transposed to [in][out]; A outputs padded to a multiple of 4
struct PpoNet {
    float *W1t, *b1, *W2t, *b2, *W3t, *b3;
    int Apad;
    __device__ __forceinline__ static int floats(int nIn, int A) { const int Ap = (A + 3) & ~3; return nIn * kPpoH + kPpoH + kPpoH * kPpoH + kPpoH + kPpoH * Ap + Ap; }
    __device__ __forceinline__ void carve(float *base, int nIn, int A)
    {
        Apad = (A + 3) & ~3;
        W1t = base; b1 = W1t + nIn * kPpoH; W2t = b1 + kPpoH; b2 = W2t + kPpoH * kPpoH; W3t = b2 + kPpoH; b3 = W3t + kPpoH * Apad;
    }
    __device__ __forceinline__ void stage(const float *w, int nIn, int A)
    {
        constexpr int H = kPpoH;
        const float *w2 = w + H * nIn + H, *w3 = w2 + H * H + H;
        for (int i = threadIdx.x; i < H * nIn; i += blockDim.x) { const int o = i / nIn, k = i - o * nIn; W1t[k * H + o] = w[i]; }
        for (int i = threadIdx.x; i < H * H; i += blockDim.x) { const int o = i / H, k = i - o * H; W2t[k * H + o] = w2[i]; }
        for (int i = threadIdx.x; i < H * Apad; i += blockDim.x) { const int k = i / Apad, o = i - k * Apad; W3t[i] = o < A ? w3[o * H + k] : 0.f; }
        for (int i = threadIdx.x; i < H; i += blockDim.x) { b1[i] = w[H * nIn + i]; b2[i] = w2[H * H + i]; }
        for (int i = threadIdx.x; i < Apad; i += blockDim.x) b3[i] = i < A ? w3[A * H + i] : 0.f;
    }
};

// acc[nt] (16 x 8 each) += sum over the warp's 32 samples r of Am[m][r] * Bn[nt*8 + n][r]; both buffers are
// element-major [element][kPpoStride].  Elements of Bn at or beyond nValid are treated as zero (stale data).
template <int NT>
__device__ __forceinline__ void outer_accumulate(float (&acc)[NT][4], const float *Am, const float *Bn, int nValid, int lane)
{
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 8) {
        uint32_t ah[4], al[4];
        tf32_split(Am[g * kPpoStride + r0 + t], ah[0], al[0]);
        tf32_split(Am[(g + 8) * kPpoStride + r0 + t], ah[1], al[1]);
        tf32_split(Am[g * kPpoStride + r0 + t + 4], ah[2], al[2]);
        tf32_split(Am[(g + 8) * kPpoStride + r0 + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int n = nt * 8 + g;
            const bool ok = n < nValid;
            const float v0 = ok ? Bn[n * kPpoStride + r0 + t] : 0.f;
            const float v1 = ok ? Bn[n * kPpoStride + r0 + t + 4] : 0.f;
            uint32_t bh0, bl0, bh1, bl1;
            tf32_split(v0, bh0, bl0);
            tf32_split(v1, bh1, bl1);
            mma_tf32(acc[nt], al, bh0, bh1);  // small terms first
            mma_tf32(acc[nt], ah, bl0, bl1);
            mma_tf32(acc[nt], ah, bh0, bh1);
        }
    }
}

// The same contraction on the FP32 pipe (same accumulator ownership as the MMA C fragment: lane (g, t) owns
// rows g, g+8 x columns 2t, 2t+1 of every 8-column block), four samples per 128-bit shared-memory load.
// Legacy mma.sync TF32 runs at the FFMA rate on sm_100a (one m16n8k8 per 32 cycles per SM sub-partition) and
// 3xTF32 needs three per product, so plain FFMA does the job in a third of the pipe time, in exact fp32.
template <int NT>
__device__ __forceinline__ void outer_accumulate_simt(float (&acc)[NT][4], const float *Am, const float *Bn, int nValid, int lane)
{
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 4) {
        const float4 a0 = *reinterpret_cast<const float4 *>(Am + g * kPpoStride + r0);
        const float4 a1 = *reinterpret_cast<const float4 *>(Am + (g + 8) * kPpoStride + r0);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int n = nt * 8 + 2 * t;
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (n < nValid) b0 = *reinterpret_cast<const float4 *>(Bn + n * kPpoStride + r0);
            if (n + 1 < nValid) b1 = *reinterpret_cast<const float4 *>(Bn + (n + 1) * kPpoStride + r0);
            acc[nt][0] = fmaf(a0.x, b0.x, acc[nt][0]); acc[nt][0] = fmaf(a0.y, b0.y, acc[nt][0]);
            acc[nt][0] = fmaf(a0.z, b0.z, acc[nt][0]); acc[nt][0] = fmaf(a0.w, b0.w, acc[nt][0]);
            acc[nt][1] = fmaf(a0.x, b1.x, acc[nt][1]); acc[nt][1] = fmaf(a0.y, b1.y, acc[nt][1]);
            acc[nt][1] = fmaf(a0.z, b1.z, acc[nt][1]); acc[nt][1] = fmaf(a0.w, b1.w, acc[nt][1]);
            acc[nt][2] = fmaf(a1.x, b0.x, acc[nt][2]); acc[nt][2] = fmaf(a1.y, b0.y, acc[nt][2]);
            acc[nt][2] = fmaf(a1.z, b0.z, acc[nt][2]); acc[nt][2] = fmaf(a1.w, b0.w, acc[nt][2]);
            acc[nt][3] = fmaf(a1.x, b1.x, acc[nt][3]); acc[nt][3] = fmaf(a1.y, b1.y, acc[nt][3]);
            acc[nt][3] = fmaf(a1.z, b1.z, acc[nt][3]); acc[nt][3] = fmaf(a1.w, b1.w, acc[nt][3]);
        }
    }
}

#ifdef MSCHED_PPO_SIMT_ACC
#define PPO_OUTER outer_accumulate_simt
#else
#define PPO_OUTER outer_accumulate
#endif

// sum over the warp's samples of one element row: lane l adds 16 samples of element l & 15 (the two
// halves are combined when the CTA partial is written)
__device__ __forceinline__ float row_half_sum(const float *buf, int lane)
{
    const float4 *p = reinterpret_cast<const float4 *>(buf + (lane & 15) * kPpoStride + (lane >> 4) * 16);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 v = p[q]; s += (v.x + v.y) + (v.z + v.w); }
    return s;
}

// The per-sample sweeps are written as RUNTIME loops over activations parked in shared memory (the buffers the
// tensor-core contraction needs anyway) instead of fully unrolled register arrays: the first version's loop body
// was 5,400 instructions (86 KB), more than the instruction cache holds for four CTAs in different phases
// (ncu: no_instruction 1.3 stalls per issue).  Buffers per warp, element-major [element][kPpoStride]:
// bufX input, bufH1 first hidden layer, bufH second hidden layer, bufD the current dL/d(pre-activation).

// forward: h1 -> bufH1, h2 -> bufH (and returned in registers for the head)
__device__ __forceinline__ void forward_hidden(const PpoNet &n, const float *bufX, float *bufH1, float *bufH, int nIn, int lane,
                                               float (&h2)[kPpoH])
{
    constexpr int H = kPpoH;
    float h1[H];
#pragma unroll
    for (int o = 0; o < H; ++o) h1[o] = n.b1[o];
#pragma unroll kPpoUnroll
    for (int k = 0; k < nIn; ++k) {
        const float xv = bufX[k * kPpoStride + lane];
        const float4 *wr = reinterpret_cast<const float4 *>(n.W1t + k * H);
#pragma unroll
        for (int o4 = 0; o4 < H / 4; ++o4) {
            const float4 wv = wr[o4];
            h1[4 * o4] = fmaf(wv.x, xv, h1[4 * o4]); h1[4 * o4 + 1] = fmaf(wv.y, xv, h1[4 * o4 + 1]);
            h1[4 * o4 + 2] = fmaf(wv.z, xv, h1[4 * o4 + 2]); h1[4 * o4 + 3] = fmaf(wv.w, xv, h1[4 * o4 + 3]);
        }
    }
    __syncwarp();  // the previous head's fragment loads of bufH1 / bufH are done
#pragma unroll
    for (int o = 0; o < H; ++o) { bufH1[o * kPpoStride + lane] = fast_tanh(h1[o]); h2[o] = n.b2[o]; }
#pragma unroll kPpoUnroll
    for (int k = 0; k < H; ++k) {
        const float xv = bufH1[k * kPpoStride + lane];  // own column: no synchronisation needed
        const float4 *wr = reinterpret_cast<const float4 *>(n.W2t + k * H);
#pragma unroll
        for (int o4 = 0; o4 < H / 4; ++o4) {
            const float4 wv = wr[o4];
            h2[4 * o4] = fmaf(wv.x, xv, h2[4 * o4]); h2[4 * o4 + 1] = fmaf(wv.y, xv, h2[4 * o4 + 1]);
            h2[4 * o4 + 2] = fmaf(wv.z, xv, h2[4 * o4 + 2]); h2[4 * o4 + 3] = fmaf(wv.w, xv, h2[4 * o4 + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < H; ++o) { h2[o] = fast_tanh(h2[o]); bufH[o * kPpoStride + lane] = h2[o]; }
}

// backward from the head's dL/dz (dz[0..AZ), already parked in bufD rows 0..AZ and contracted with bufH for dW3):
// da2 = (W3^T dz) * (1 - h2^2) -> bufD, dW2 += da2 (x) h1, da1 = (W2^T da2) * (1 - h1^2) -> bufD, dW1 += da1 (x) x
template <int NT1, int AZ>
__device__ __forceinline__ void backward_hidden(const PpoNet &n, const float (&dz)[AZ], float *bufX, float *bufH1, float *bufH, float *bufD,
                                                int nIn, int lane, float (&acc1)[NT1][4], float (&acc2)[2][4], float &db1, float &db2)
{
    constexpr int H = kPpoH;
    __syncwarp();  // the dW3 fragment loads of bufD (dz) are done
#pragma unroll kPpoUnroll
    for (int k = 0; k < H; ++k) {
        float s = 0.f;
        if (AZ == 1) {
            s = n.W3t[k * n.Apad] * dz[0];
        } else {
            const float4 *wr = reinterpret_cast<const float4 *>(n.W3t + k * n.Apad);
#pragma unroll
            for (int o4 = 0; o4 < AZ / 4; ++o4) {
                if (4 * o4 < n.Apad) {
                    const float4 wv = wr[o4];
                    s = fmaf(wv.x, dz[4 * o4], s); s = fmaf(wv.y, dz[4 * o4 + 1], s);
                    s = fmaf(wv.z, dz[4 * o4 + 2], s); s = fmaf(wv.w, dz[4 * o4 + 3], s);
                }
            }
        }
        const float h = bufH[k * kPpoStride + lane];
        bufD[k * kPpoStride + lane] = s * fmaf(-h, h, 1.f);
    }
    __syncwarp();
    PPO_OUTER<2>(acc2, bufD, bufH1, H, lane);  // dW2[k_out][i_in]
    db2 += row_half_sum(bufD, lane);
    float da2[H];
#pragma unroll
    for (int k = 0; k < H; ++k) da2[k] = bufD[k * kPpoStride + lane];
    __syncwarp();  // the dW2 fragment loads of bufD are done
#pragma unroll kPpoUnroll
    for (int i = 0; i < H; ++i) {
        const float4 *wr = reinterpret_cast<const float4 *>(n.W2t + i * H);
        float s = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 wv = wr[k4];
            s = fmaf(wv.x, da2[4 * k4], s); s = fmaf(wv.y, da2[4 * k4 + 1], s);
            s = fmaf(wv.z, da2[4 * k4 + 2], s); s = fmaf(wv.w, da2[4 * k4 + 3], s);
        }
        const float h = bufH1[i * kPpoStride + lane];
        bufD[i * kPpoStride + lane] = s * fmaf(-h, h, 1.f);
    }
    __syncwarp();
    PPO_OUTER<NT1>(acc1, bufD, bufX, nIn, lane);  // dW1[k_out][i_in]
    db1 += row_half_sum(bufD, lane);
}

// write one net's accumulators of this warp into the CTA sum (shared memory, torch parameter order)
template <int NT1, int NT3>
__device__ __forceinline__ void scatter_accumulators(float *red, int nIn, int A, int lane, const float (&acc1)[NT1][4],
                                                     const float (&acc2)[2][4], const float (&acc3)[NT3][4], float db1, float db2, float db3)
{
    constexpr int H = kPpoH;
    const int g = lane >> 2, t = lane & 3;
    const int off2 = H * nIn + H, off3 = off2 + H * H + H;
#pragma unroll
    for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = g + ((c & 2) ? 8 : 0), i = nt * 8 + 2 * t + (c & 1);
            if (i < nIn) red[k * nIn + i] += acc1[nt][c];
        }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = g + ((c & 2) ? 8 : 0), i = nt * 8 + 2 * t + (c & 1);
            red[off2 + k * H + i] += acc2[nt][c];
        }
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = g + ((c & 2) ? 8 : 0), j = nt * 8 + 2 * t + (c & 1);
            if (j < A) red[off3 + j * H + k] += acc3[nt][c];
        }
    // the two 16-sample halves of the bias sums
    db1 += __shfl_xor_sync(0xffffffffu, db1, 16);
    db2 += __shfl_xor_sync(0xffffffffu, db2, 16);
    db3 += __shfl_xor_sync(0xffffffffu, db3, 16);
    if (lane < H) {
        red[H * nIn + lane] += db1;
        red[off2 + H * H + lane] += db2;
        if (lane < A) red[off3 + A * H + lane] += db3;
    }
}

// NT1 = ceil(nIn / 8) input tiles, NT3 = ceil(A / 8) action tiles
#ifndef MSCHED_PPO_MINB
#define MSCHED_PPO_MINB 4
#endif

template <int NT1, int NT3>
__global__ void __launch_bounds__(128, MSCHED_PPO_MINB) ppo_grad_kernel(const PpoArgs a)
{
    constexpr int H = kPpoH, AP = NT3 * 8;
    extern __shared__ __align__(16) float sm[];
    const int nIn = a.nIn, A = a.A;
    const int sel = blockIdx.y;
    const int net = a.netIds[sel];
    const int pcA = ppo_param_count(nIn, A), pcC = ppo_param_count(nIn, 1);
    const int P = pcA + pcC + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    PpoNet act, cri;
    float *p = sm;
    act.carve(p, nIn, A); p += (PpoNet::floats(nIn, A) + 3) & ~3;
    cri.carve(p, nIn, 1); p += (PpoNet::floats(nIn, 1) + 3) & ~3;
    float *red = p;  // the CTA sum aliases the warps' buffers: it is only used after the sample loop
    float *bufX = p + warp * (NT1 * 8 + 3 * H) * kPpoStride;
    float *bufH1 = bufX + NT1 * 8 * kPpoStride, *bufH = bufH1 + H * kPpoStride, *bufD = bufH + H * kPpoStride;
    act.stage(a.actorW + (size_t)net * pcA, nIn, A);
    cri.stage(a.criticW + (size_t)net * pcC, nIn, 1);
    __syncthreads();

    float a1[NT1][4], a2[2][4], a3[NT3][4], c1[NT1][4], c2[2][4], c3[1][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int nt = 0; nt < NT1; ++nt) a1[nt][c] = c1[nt][c] = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) a3[nt][c] = 0.f;
        a2[0][c] = a2[1][c] = c2[0][c] = c2[1][c] = c3[0][c] = 0.f;
    }
    float adb1 = 0.f, adb2 = 0.f, adb3 = 0.f, cdb1 = 0.f, cdb2 = 0.f, cdb3 = 0.f;
    float sSurr = 0.f, sMse = 0.f, sEnt = 0.f;

    const long long total = a.nTb * a.m;
    const long long nTiles = (total + 127) / 128;
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const long long q = tile * 128 + threadIdx.x;
        const bool live = q < total;
        long long tb = 0;
        int u = 0;
        if (live) {
            tb = q / a.m;
            u = a.unitIds[(size_t)sel * a.m + (int)(q - tb * a.m)];
        }
        const long long r = tb * a.U + u;
        const int16_t *xr = a.x + tb * a.xTbStride + (long long)u * a.xUnitStride;
#ifdef MSCHED_PPO_PREFETCH  // measured: no gain once four CTAs per SM hide the latency (1.47 ms with, 1.43 ms without)
        {   // the next tile's rows start their way from HBM now (strided 2-byte loads are latency-bound)
            const long long q2 = q + (long long)gridDim.x * 128;
            if (q2 < total) {
                const long long tb2 = q2 / a.m;
                const int u2 = a.unitIds[(size_t)sel * a.m + (int)(q2 - tb2 * a.m)];
                const int16_t *x2 = a.x + tb2 * a.xTbStride + (long long)u2 * a.xUnitStride;
                prefetch_l1(x2);
                prefetch_l1(x2 + nIn - 1);
                const long long r2 = tb2 * a.U + u2;
                prefetch_l1(a.ret + r2);
                prefetch_l1(a.action + r2);
                prefetch_l1(a.logpOld + r2);
            }
        }
#endif
        __syncwarp();  // last tile's fragment loads of bufX are done
        for (int k = 0; k < NT1 * 8; ++k) bufX[k * kPpoStride + lane] = (live && k < nIn) ? (float)xr[k] : 0.f;
        const float G = live ? a.ret[r] : 0.f;
        const int aSel = live ? a.action[r] : 0;
        const float lpOld = live ? a.logpOld[r] : 0.f;
        const float lv = live ? 1.f : 0.f;

        // ---- critic: V(s), dL/dV = 2*c_val*(V - G) ----
        float V;
        {
            float h2[H];
            forward_hidden(cri, bufX, bufH1, bufH, nIn, lane, h2);
            V = cri.b3[0];
#pragma unroll
            for (int k = 0; k < H; ++k) V = fmaf(cri.W3t[k * cri.Apad], h2[k], V);
            float dv[1] = {2.f * a.valCoef * (V - G) * lv};
            sMse += (V - G) * (V - G) * lv;
            bufD[lane] = dv[0];  // (the forward's barrier already separates this from the previous fragment loads)
            __syncwarp();
            PPO_OUTER<1>(c3, bufH, bufD, 1, lane);  // dW3[0][k]: m = k, n = 0
            cdb3 += row_half_sum(bufD, lane);       // only lanes 0 and 16 hold element 0
            backward_hidden<NT1, 1>(cri, dv, bufX, bufH1, bufH, bufD, nIn, lane, c1, c2, cdb1, cdb2);
        }
        // ---- actor: log-softmax, clipped surrogate and entropy gradients ----
        {
            float h2[H], z[AP];
            forward_hidden(act, bufX, bufH1, bufH, nIn, lane, h2);
#pragma unroll
            for (int o4 = 0; o4 < AP / 4; ++o4) {
                if (4 * o4 < act.Apad) {
                    float4 s = *reinterpret_cast<const float4 *>(act.b3 + 4 * o4);
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        const float4 wv = *reinterpret_cast<const float4 *>(act.W3t + k * act.Apad + 4 * o4);
                        s.x = fmaf(wv.x, h2[k], s.x); s.y = fmaf(wv.y, h2[k], s.y);
                        s.z = fmaf(wv.z, h2[k], s.z); s.w = fmaf(wv.w, h2[k], s.w);
                    }
                    z[4 * o4] = s.x; z[4 * o4 + 1] = s.y; z[4 * o4 + 2] = s.z; z[4 * o4 + 3] = s.w;
                } else {
                    z[4 * o4] = z[4 * o4 + 1] = z[4 * o4 + 2] = z[4 * o4 + 3] = -INFINITY;
                }
            }
#pragma unroll
            for (int o = 0; o < AP; ++o) z[o] = o < A ? z[o] : -INFINITY;
            float mx = z[0];
#pragma unroll
            for (int o = 1; o < AP; ++o) mx = fmaxf(mx, z[o]);
            float se = 0.f;
#pragma unroll
            for (int o = 0; o < AP; ++o) se += __expf(z[o] - mx);
            const float lse = mx + logf(se);
            float ent = 0.f, lpa = 0.f;
            float pr[AP];
#pragma unroll
            for (int o = 0; o < AP; ++o) {
                const float lp = z[o] - lse;      // -inf for padded actions
                pr[o] = __expf(lp);               // 0 for padded actions
                ent -= o < A ? pr[o] * lp : 0.f;
                lpa = o == aSel ? lp : lpa;
                z[o] = lp;
            }
            const float ratio = __expf(lpa - lpOld);
            const float adv = G - V;
            const float lo = 1.f - a.epsClip, hi = 1.f + a.epsClip;
            const float s1 = ratio * adv, s2 = fminf(fmaxf(ratio, lo), hi) * adv;
            const float d = (s1 < s2 || (ratio >= lo && ratio <= hi)) ? adv : 0.f;
            sSurr -= fminf(s1, s2) * lv;
            sEnt += ent * lv;
            const float cr = -d * ratio;
            float dz[AP];
#pragma unroll
            for (int o = 0; o < AP; ++o) {
                const float ind = o == aSel ? 1.f : 0.f;
                dz[o] = o < A ? (cr * (ind - pr[o]) + a.entCoef * pr[o] * (z[o] + ent)) * lv : 0.f;
            }
#pragma unroll
            for (int o = 0; o < AP; ++o) bufD[o * kPpoStride + lane] = dz[o];
            __syncwarp();
            PPO_OUTER<NT3>(a3, bufH, bufD, A, lane);  // dW3[j][k]: m = k, n = j
            adb3 += row_half_sum(bufD, lane);
            backward_hidden<NT1, AP>(act, dz, bufX, bufH1, bufH, bufD, nIn, lane, a1, a2, adb1, adb2);
        }
    }

    // ---- CTA partial: the four warps add their accumulators in turn (fixed order) ----
    __syncthreads();  // every warp is done with its buffers, which `red` overlays
    for (int i = threadIdx.x; i < P; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
            scatter_accumulators<NT1, NT3>(red, nIn, A, lane, a1, a2, a3, adb1, adb2, adb3);
            scatter_accumulators<NT1, 1>(red + pcA, nIn, 1, lane, c1, c2, c3, cdb1, cdb2, cdb3);
            float s0 = sSurr, s1 = sMse, s2 = sEnt;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) { red[pcA + pcC] += s0; red[pcA + pcC + 1] += s1; red[pcA + pcC + 2] += s2; }
        }
        __syncthreads();
    }
    float *out = a.partial + ((size_t)sel * gridDim.x + blockIdx.x) * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) out[i] = red[i];
}

// grad[net][i] = (1/M) * sum over CTAs (fixed order) of the partial sums; stats[sel] = means and M
__global__ void ppo_reduce_kernel(const float *__restrict__ partial, int nCta, int pcA, int pcC, const int32_t *__restrict__ netIds,
                                  float invM, float M, float *__restrict__ gradA, float *__restrict__ gradC, float *__restrict__ stats)
{
    const int sel = blockIdx.y;
    const int P = pcA + pcC + 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float *p = partial + (size_t)sel * nCta * P + i;
    float s = 0.f;
    for (int c = 0; c < nCta; ++c) s += p[(size_t)c * P];
    const int net = netIds[sel];
    if (i < pcA) gradA[(size_t)net * pcA + i] = s * invM;
    else if (i < pcA + pcC) gradC[(size_t)net * pcC + (i - pcA)] = s * invM;
    else if (stats) stats[sel * 4 + (i - pcA - pcC)] = (i == P - 1) ? M : s * invM;
}

// torch.optim.Adam's single-tensor step (no weight decay, no amsgrad), same operation order:
// m.lerp_(g, 1-b1); v = v*b2 + (1-b2)*g*g; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            long long n, float stepSize, float oneMinusB1, float b2, float oneMinusB2, float bc2Sqrt, float eps)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = __fmaf_rn(gi - m[i], oneMinusB1, m[i]);
    const float vi = __fmaf_rn(oneMinusB2 * gi, gi, __fmul_rn(v[i], b2));
    m[i] = mi;
    v[i] = vi;
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2Sqrt), eps);
    p[i] = __fmaf_rn(-stepSize, __fdiv_rn(mi, denom), p[i]);
}

}  // namespace msched
