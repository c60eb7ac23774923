// msched_dqn.cu -- optimize_model of the DQN learner (src/DQNmodules.py:97-154) on the device: the gradient of the
// SmoothL1 loss between Q_policy(s)[a] and gamma * max_a' Q_target(s') + r for every Q-net of a group
// (DQNEntity: Linear(in,16)-Tanh-Linear(16,A), src/DQNmodules.py:41-46), clamped to [-1, 1] like the reference
// does before optimizer.step().  One CTA per net.  The batch is walked in chunks of 128 transitions: first one
// thread per transition runs both forwards and parks the hidden activations, dL/dz of the hidden layer and the
// Huber slope in shared memory, then one thread per PARAMETER sums its gradient over the chunk in sample order --
// no atomics, so the gradient is bit-reproducible.  msched_adam_step applies torch.optim.Adam.
#include <cstring>

#include "abi_common.h"
#include "msched_common.cuh"

namespace msched {

constexpr int kDqnH = 16, kDqnChunk = 128, kDqnMaxR = 26;  // up to 26 * 128 = 3,328 parameters per net

struct DqnArgs {
    const float *policy, *target;
    int nIn, A, nNets, batch;
    const int16_t *state, *nextState;  // [batch][nNets][nIn]
    const int32_t *action;             // [batch][nNets]
    const float *reward;               // [batch][nNets]
    float gamma;
    float *grad;                       // [nNets][pc]
    float *loss;                       // [nNets] or null
};

__global__ void __launch_bounds__(kDqnChunk) dqn_grad_kernel(const DqnArgs a)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int H = kDqnH;
    const int nIn = a.nIn, A = a.A, net = blockIdx.x, tid = threadIdx.x;
    const int pc = H * nIn + H + A * H + A;
    float *wp = sm, *wt = wp + pc;                  // policy / target parameters, torch layout
    float *sh = wt + pc;                            // h [chunk][H]
    float *sdz = sh + kDqnChunk * H;                // dL/dz1 [chunk][H]
    float *sg = sdz + kDqnChunk * H;                // Huber slope / batch [chunk]
    int *sa = reinterpret_cast<int *>(sg + kDqnChunk);  // action [chunk]
    int16_t *sx = reinterpret_cast<int16_t *>(sa + kDqnChunk);  // state rows [chunk][nIn]
    __shared__ float redLoss[kDqnChunk / 32];
    for (int i = tid; i < pc; i += kDqnChunk) { wp[i] = a.policy[(size_t)net * pc + i]; wt[i] = a.target[(size_t)net * pc + i]; }
    __syncthreads();
    const float *W1 = wp, *b1 = wp + H * nIn, *W2 = b1 + H, *b2 = W2 + A * H;
    const float *T1 = wt, *tb1 = wt + H * nIn, *T2 = tb1 + H, *tb2 = T2 + A * H;
    float acc[kDqnMaxR];
#pragma unroll
    for (int r = 0; r < kDqnMaxR; ++r) acc[r] = 0.f;
    float lossSum = 0.f;
    const float invB = 1.f / (float)a.batch;
    for (int c0 = 0; c0 < a.batch; c0 += kDqnChunk) {
        const int s = c0 + tid;
        const bool live = s < a.batch;
        // ---- phase 1: one thread per transition ----
        {
            float h[H], ht[H];
#pragma unroll
            for (int k = 0; k < H; ++k) { h[k] = b1[k]; ht[k] = tb1[k]; }
            const int16_t *xs = a.state + ((size_t)(live ? s : 0) * a.nNets + net) * nIn;
            const int16_t *xn = a.nextState + ((size_t)(live ? s : 0) * a.nNets + net) * nIn;
            for (int i = 0; i < nIn; ++i) {
                const int16_t xi = xs[i];
                sx[tid * nIn + i] = live ? xi : (int16_t)0;
                const float xv = (float)xi, xw = (float)xn[i];
#pragma unroll
                for (int k = 0; k < H; ++k) { h[k] = fmaf(W1[k * nIn + i], xv, h[k]); ht[k] = fmaf(T1[k * nIn + i], xw, ht[k]); }
            }
#pragma unroll
            for (int k = 0; k < H; ++k) { h[k] = tanhf(h[k]); ht[k] = tanhf(ht[k]); }
            const int act = live ? a.action[(size_t)s * a.nNets + net] : 0;
            float qa = 0.f, best = -INFINITY;
            for (int o = 0; o < A; ++o) {
                float q = b2[o], qt = tb2[o];
#pragma unroll
                for (int k = 0; k < H; ++k) { q = fmaf(W2[o * H + k], h[k], q); qt = fmaf(T2[o * H + k], ht[k], qt); }
                if (o == act) qa = q;
                best = fmaxf(best, qt);
            }
            const float expected = fmaf(best, a.gamma, live ? a.reward[(size_t)s * a.nNets + net] : 0.f);
            const float d = qa - expected;
            const float g = live ? fminf(fmaxf(d, -1.f), 1.f) * invB : 0.f;  // d/dq of the batch-mean SmoothL1Loss (beta 1)
            if (live) lossSum += fabsf(d) < 1.f ? 0.5f * d * d : fabsf(d) - 0.5f;
            sg[tid] = g;
            sa[tid] = act;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                sh[tid * H + k] = h[k];
                sdz[tid * H + k] = g * W2[act * H + k] * (1.f - h[k] * h[k]);
            }
        }
        __syncthreads();
        // ---- phase 2: one thread per parameter, samples in order ----
        const int n = min(kDqnChunk, a.batch - c0);
#pragma unroll
        for (int r = 0; r < kDqnMaxR; ++r) {
            const int p = r * kDqnChunk + tid;
            if (p >= pc) break;
            float v = acc[r];
            if (p < H * nIn) {
                const int k = p / nIn, i = p - k * nIn;
                for (int t = 0; t < n; ++t) v = fmaf(sdz[t * H + k], (float)sx[t * nIn + i], v);
            } else if (p < H * nIn + H) {
                const int k = p - H * nIn;
                for (int t = 0; t < n; ++t) v += sdz[t * H + k];
            } else if (p < H * nIn + H + A * H) {
                const int q = p - H * nIn - H, o = q / H, k = q - o * H;
                for (int t = 0; t < n; ++t) v = sa[t] == o ? fmaf(sg[t], sh[t * H + k], v) : v;
            } else {
                const int o = p - H * nIn - H - A * H;
                for (int t = 0; t < n; ++t) v += sa[t] == o ? sg[t] : 0.f;
            }
            acc[r] = v;
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < kDqnMaxR; ++r) {
        const int p = r * kDqnChunk + tid;
        if (p < pc) a.grad[(size_t)net * pc + p] = fminf(fmaxf(acc[r], -1.f), 1.f);  // param.grad.data.clamp_(-1, 1)
    }
    if (a.loss) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lossSum += __shfl_xor_sync(0xffffffffu, lossSum, o);
        if ((tid & 31) == 0) redLoss[tid >> 5] = lossSum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < kDqnChunk / 32; ++w) t += redLoss[w];
            a.loss[net] = t * invB;
        }
    }
}

}  // namespace msched

using namespace msched;

extern "C" int msched_dqn_grad(const MschedDqnBatch *b, void *stream)
{
    if (!b || !b->policy || !b->target || !b->state || !b->next_state || !b->action || !b->reward || !b->grad)
        return fail(MSCHED_E_ARG, "null DQN batch field");
    if (b->n_hidden != kDqnH) return fail(MSCHED_E_ARG, "the DQN nets have 16 hidden neurons (src/DQNmodules.py:41-46)");
    if (b->n_in < 1 || b->n_actions < 1 || b->n_nets < 1 || b->batch < 1) return fail(MSCHED_E_ARG, "bad n_in/n_actions/n_nets/batch");
    const int pc = kDqnH * b->n_in + kDqnH + b->n_actions * kDqnH + b->n_actions;
    if (pc > kDqnMaxR * kDqnChunk) return fail(MSCHED_E_ARG, "Q-net too large for the DQN gradient kernel (3,328 parameters)");
    DqnArgs a;
    a.policy = b->policy; a.target = b->target; a.nIn = b->n_in; a.A = b->n_actions; a.nNets = b->n_nets; a.batch = b->batch;
    a.state = b->state; a.nextState = b->next_state; a.action = b->action; a.reward = b->reward; a.gamma = b->gamma;
    a.grad = b->grad; a.loss = b->loss;
    size_t smem = sizeof(float) * ((size_t)2 * pc + 2 * kDqnChunk * kDqnH + kDqnChunk) + sizeof(int) * kDqnChunk +
                  sizeof(int16_t) * (size_t)kDqnChunk * b->n_in;
    smem = (smem + 15) & ~(size_t)15;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(dqn_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dqn_grad_kernel<<<b->n_nets, kDqnChunk, smem, static_cast<cudaStream_t>(stream)>>>(a);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}
