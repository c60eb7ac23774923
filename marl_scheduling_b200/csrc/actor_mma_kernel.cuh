// actor_mma_kernel.cuh -- ActorCritic.act (src/PPOmodules.py:53-63) for the 16-wide nets of the divided /
// shared agents on the tensor cores, warp-level.
//
// A 16-wide net is too small for a tcgen05 tile (three UMMA round trips through shared memory, an
// mbarrier and a tcgen05.ld per 128 rows cost more than the 608 FFMA per row they replace -- measured,
// DESIGN.md section 5), but it maps exactly onto warp-level m16n8k8 TF32 MMAs with NO shared memory at all:
//   * a warp owns 32 rows (two 16-row M tiles); lane (g, t) = (lane / 4, lane % 4) sees rows g, g + 8;
//   * the WEIGHTS are the B operands and live in registers for the whole persistent loop, already split
//     hi + lo (3xTF32: hi*hi + hi*lo + lo*hi, fp32 accuracy) -- 16 + 16 + 8*NT3 registers;
//   * the contraction index may be enumerated in any order, so operand slot k = t is bound to column
//     2t and slot k = t + 4 to column 2t + 1 of each 8-column block: with that binding the C fragment of
//     one layer (row g, columns 2t and 2t + 1) IS the A fragment of the next layer -- activations go
//     from layer to layer in registers, no shuffles, no shared memory; bias is the accumulator's
//     initial value;
//   * softmax, the inverse-CDF categorical sample and the log-prob run on the C fragments of the last
//     layer: a row's logits sit in the 4 lanes of a quad, reductions and the prefix sum are quad
//     shuffles, and lane t of the quad finishes row t of the quad's four rows (one Philox call each).
// Per 32 rows: 60-72 MMAs + ~560 other warp instructions instead of ~1,850 for the fp32 SIMT kernel.
#pragma once
#include "policy_common.cuh"

namespace msched {

template <int NT3>
__device__ __forceinline__ void quad_row_epilogue(const ActorArgs &a, float (&lg)[NT3][2], int A, int t, float u,
                                                  int &actOut, float &lpOut, float (&pOut)[NT3][2])
{
    // base-2 logits of one row: this lane holds columns 8*nt + 2t + j; padded columns are -inf
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) mx = fmaxf(mx, fmaxf(lg[nt][0], lg[nt][1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float s = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) { lg[nt][0] = ex2_approx(lg[nt][0] - mx); lg[nt][1] = ex2_approx(lg[nt][1] - mx); s += lg[nt][0] + lg[nt][1]; }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float inv = 1.f / s;
    // Categorical(probs) renormalises by the sum of the softmax output
    float tot = 0.f, pair[NT3];
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) {
        lg[nt][0] *= inv; lg[nt][1] *= inv;
        pOut[nt][0] = lg[nt][0]; pOut[nt][1] = lg[nt][1];
        pair[nt] = lg[nt][0] + lg[nt][1];
        tot += pair[nt];
    }
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
    // inverse CDF over the columns in order: prefix sum inside the quad per 8-column block
    const float thr = u * tot;
    float base = 0.f;
    int act = 0x7fffffff;
    float pa = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) {
        float inc = pair[nt];
        float up = __shfl_up_sync(0xffffffffu, inc, 1, 4);
        inc += t >= 1 ? up : 0.f;
        up = __shfl_up_sync(0xffffffffu, inc, 2, 4);
        inc += t >= 2 ? up : 0.f;
        const float c0 = base + (inc - pair[nt]) + lg[nt][0], c1 = c0 + lg[nt][1];
        if (act == 0x7fffffff) {
            if (c0 > thr) { act = 8 * nt + 2 * t; pa = lg[nt][0]; }
            else if (c1 > thr) { act = 8 * nt + 2 * t + 1; pa = lg[nt][1]; }
        }
        base += __shfl_sync(0xffffffffu, inc, 3, 4);
    }
    // first column over the quad whose running sum exceeds the threshold (padded columns add 0: never first)
    int best = act;
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
    if (best == 0x7fffffff) {  // u * total rounded up to the total: the last action (its owner lane supplies pa)
        best = A - 1;
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) pa = (8 * nt + 2 * t + j == A - 1) ? lg[nt][j] : pa;
    }
    const int owner = (best & 7) >> 1;
    pa = __shfl_sync(0xffffffffu, pa, owner, 4);
    actOut = best;
    const float eps = 1.1920928955078125e-07f;
    float pn = pa / tot;
    pn = fminf(fmaxf(pn, eps), 1.f - eps);
    lpOut = logf(pn);
}

// NT1 = input tiles of 8 (n_in <= 8*NT1), NT3 = action tiles of 8 (A <= 8*NT3); 16 hidden neurons
template <int NT1, int NT3>
__global__ void __launch_bounds__(128) actor_forward_mma(const ActorArgs a)
{
    constexpr int H = 16;
    const int nIn = a.nIn, A = a.nActions;
    const int unit = blockIdx.y;
    const int net = (unit / a.unitDiv) % a.nNets;
    const int pc = H * nIn + H + H * H + H + A * H + A;
    const float *w = a.weights + (size_t)net * pc;
    const float *w2 = w + H * nIn + H, *w3 = w2 + H * H + H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

    // B fragments (weights, torch [out][in]): slot k = t <-> input column 8*ks + 2t, slot t + 4 <-> 8*ks + 2t + 1
    uint32_t w1h[NT1][2][2], w1l[NT1][2][2], w2h[2][2][2], w2l[2][2][2], w3h[2][NT3][2], w3l[2][NT3][2];
    float bia1[2][2], bia2[2][2], bia3[NT3][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const int n = nt * 8 + g;
#pragma unroll
        for (int ks = 0; ks < NT1; ++ks)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = 8 * ks + 2 * t + j;
                tf32_split(k < nIn ? w[n * nIn + k] : 0.f, w1h[ks][nt][j], w1l[ks][nt][j]);
            }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int j = 0; j < 2; ++j) tf32_split(w2[n * H + 8 * ks + 2 * t + j], w2h[ks][nt][j], w2l[ks][nt][j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) { bia1[nt][j] = w[H * nIn + nt * 8 + 2 * t + j]; bia2[nt][j] = w2[H * H + nt * 8 + 2 * t + j]; }
    }
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) {
        const int n = nt * 8 + g;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int j = 0; j < 2; ++j)  // base-2 logits: log2(e) folded into the last layer
                tf32_split(n < A ? w3[n * H + 8 * ks + 2 * t + j] * kLog2e : 0.f, w3h[ks][nt][j], w3l[ks][nt][j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) { const int c = nt * 8 + 2 * t + j; bia3[nt][j] = c < A ? w3[A * H + c] * kLog2e : -INFINITY; }
    }

    const int nTiles = (a.nEnvs + 127) / 128;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int env0 = tile * 128 + warp * 32;
        if (env0 >= a.nEnvs) continue;  // warp-uniform
        // Philox draw of row #t of the quad's four rows (rows g, g+8, g+16, g+24 of the warp tile)
        const int myEnv = env0 + g + 8 * t;
        const bool myLive = myEnv < a.nEnvs;
        const long long myRow = (long long)myEnv * a.units + unit;
        float uMine = 0.f;
        if (myLive) {
            if (a.uOverride) {
                uMine = a.uOverride[myRow];
            } else {
                const unsigned long long gr = (unsigned long long)(a.rowOffset + myRow);
                const unsigned long long stp = a.stepDev ? *a.stepDev : a.step;
                uint32_t x4[4];
                philox4x32_10((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)stp, (kStreamPolicy << 28) | (uint32_t)((stp >> 32) & 0x0fffffffu),
                              (uint32_t)a.seed, (uint32_t)(a.seed >> 32), x4);
                uMine = (float)(x4[0] >> 8) * (1.0f / 16777216.0f);
            }
        }
        int actMine = 0, gselMine = -1;
        float lpMine = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            // ---- layer 1: A fragments straight from the observation rows ----
            float c[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { c[nt][0] = c[nt][2] = bia1[nt][0]; c[nt][1] = c[nt][3] = bia1[nt][1]; }
            int gsel[2] = {-1, -1};
            {
                const int16_t *xr[2];
                bool live[2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int env = env0 + mt * 16 + hh * 8 + g;
                    live[hh] = env < a.nEnvs;
                    xr[hh] = a.x + (size_t)(live[hh] ? env : 0) * a.envStride + (size_t)unit * a.unitStride;
                    if (a.gatherCore) gsel[hh] = live[hh] ? a.gatherCore[(long long)env * a.units + unit] : 0;
                }
#pragma unroll
                for (int ks = 0; ks < NT1; ++ks) {
                    float xv[4];  // a0 (row g, slot t), a1 (row g+8, slot t), a2 (row g, slot t+4), a3 (row g+8, slot t+4)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int k = 8 * ks + 2 * t + j;
                            float v = 0.f;
                            if (k < nIn && live[hh]) {
                                if (a.gatherCore) {
                                    // FreePriceOfferPPO.selectAction: [core prio, core rem, slot prio, slot rem] of the chosen
                                    // core, [-5]*4 for core action 0 (src/PPOmodules.py:312-332, quirk Q1)
                                    const int gs = gsel[hh];
                                    if (gs <= 0 || gs > a.nCores) v = -5.f;
                                    else v = (float)xr[hh][k < 2 ? 2 * gs + k : 2 * a.nCores + (k - 2)];
                                } else {
                                    v = (float)xr[hh][k];
                                }
                            }
                            xv[hh + 2 * j] = v;
                            if (a.xUsed && k < nIn && live[hh])
                                a.xUsed[((long long)(env0 + mt * 16 + hh * 8 + g) * a.units + unit) * nIn + k] = (int16_t)v;
                        }
                    uint32_t ah[4], al[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) tf32_split(xv[q], ah[q], al[q]);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        mma_tf32(c[nt], al, w1h[ks][nt][0], w1h[ks][nt][1]);
                        mma_tf32(c[nt], ah, w1l[ks][nt][0], w1l[ks][nt][1]);
                        mma_tf32(c[nt], ah, w1h[ks][nt][0], w1h[ks][nt][1]);
                    }
                }
            }
            // ---- layer 2: the C fragment (row g: cols 2t, 2t+1; row g+8: same) is the next A fragment ----
            float d[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { d[nt][0] = d[nt][2] = bia2[nt][0]; d[nt][1] = d[nt][3] = bia2[nt][1]; }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t ah[4], al[4];
                tf32_split(fast_tanh(c[ks][0]), ah[0], al[0]);  // a0: row g,   slot t   <- col 2t
                tf32_split(fast_tanh(c[ks][2]), ah[1], al[1]);  // a1: row g+8, slot t
                tf32_split(fast_tanh(c[ks][1]), ah[2], al[2]);  // a2: row g,   slot t+4 <- col 2t+1
                tf32_split(fast_tanh(c[ks][3]), ah[3], al[3]);  // a3: row g+8, slot t+4
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    mma_tf32(d[nt], al, w2h[ks][nt][0], w2h[ks][nt][1]);
                    mma_tf32(d[nt], ah, w2l[ks][nt][0], w2l[ks][nt][1]);
                    mma_tf32(d[nt], ah, w2h[ks][nt][0], w2h[ks][nt][1]);
                }
            }
            // ---- layer 3: base-2 logits ----
            float z[NT3][4];
#pragma unroll
            for (int nt = 0; nt < NT3; ++nt) { z[nt][0] = z[nt][2] = bia3[nt][0]; z[nt][1] = z[nt][3] = bia3[nt][1]; }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t ah[4], al[4];
                tf32_split(fast_tanh(d[ks][0]), ah[0], al[0]);
                tf32_split(fast_tanh(d[ks][2]), ah[1], al[1]);
                tf32_split(fast_tanh(d[ks][1]), ah[2], al[2]);
                tf32_split(fast_tanh(d[ks][3]), ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < NT3; ++nt) {
                    mma_tf32(z[nt], al, w3h[ks][nt][0], w3h[ks][nt][1]);
                    mma_tf32(z[nt], ah, w3l[ks][nt][0], w3l[ks][nt][1]);
                    mma_tf32(z[nt], ah, w3h[ks][nt][0], w3h[ks][nt][1]);
                }
            }
            // ---- softmax / sample / log-prob of the two rows (g and g+8 of this M tile) in the quad ----
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float lg[NT3][2], pr[NT3][2];
#pragma unroll
                for (int nt = 0; nt < NT3; ++nt) { lg[nt][0] = z[nt][2 * hh]; lg[nt][1] = z[nt][2 * hh + 1]; }
                const int rq = mt * 2 + hh;  // which of the quad's four rows
                const float u = __shfl_sync(0xffffffffu, uMine, (lane & ~3) | rq);
                int act;
                float lp;
                quad_row_epilogue<NT3>(a, lg, A, t, u, act, lp, pr);
                if (t == rq) { actMine = act; lpMine = lp; gselMine = gsel[hh]; }
                if (a.probs) {
                    const int env = env0 + mt * 16 + hh * 8 + g;
                    if (env < a.nEnvs) {
                        float *po = a.probs + ((size_t)env * a.units + unit) * A;
#pragma unroll
                        for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int col = 8 * nt + 2 * t + j;
                                if (col < A) po[col] = pr[nt][j];
                            }
                    }
                }
            }
        }
        if (myLive) {
            if (a.action) a.action[myRow] = actMine;
            if (a.actionRec) a.actionRec[(size_t)myEnv * a.actionRecStride + unit] = (int16_t)(gselMine == 0 ? -5 : actMine);
            if (a.logprob) a.logprob[myRow] = lpMine;
        }
    }
}

}  // namespace msched
