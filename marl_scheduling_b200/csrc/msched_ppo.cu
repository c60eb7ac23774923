// msched_ppo.cu -- PPO.update on the device (include/msched.h: msched_ppo_grad, msched_adam_step).
// Its own translation unit: the gradient kernel is instantiated per input-tile / action-tile count and is
// the slowest thing to compile in the library.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "abi_common.h"
#include "msched_common.cuh"
#include "ppo_update_kernel.cuh"

using namespace msched;

namespace {

// CTAs per selected net: the grid is persistent, one wave of about 4 CTAs per SM shared by the nets
int ppo_grid_x(const MschedPpoBatch *b)
{
    const long long total = (long long)b->n_tb * b->units_per_net;
    long long tiles = (total + 127) / 128;
    long long gx = (148 * MSCHED_PPO_MINB) / b->n_sel;  // one resident wave: MSCHED_PPO_MINB CTAs per SM
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    return (int)gx;
}

int ppo_check(const MschedPpoBatch *b)
{
    if (!b) return fail(MSCHED_E_ARG, "null batch");
    if (b->n_hidden != 16 || b->n_in < 1 || b->n_in > 64 || b->n_actions < 1 || b->n_actions > 16)
        return fail(MSCHED_E_ARG, "ppo_grad: 16 hidden neurons, n_in 1..64, n_actions 1..16 (other shapes: autograd)");
    if (b->n_nets < 1 || b->n_sel < 1 || b->n_sel > 65535 || b->units_per_net < 1 || b->units < 1 || b->n_tb < 1)
        return fail(MSCHED_E_ARG, "ppo_grad: bad n_nets/n_sel/units_per_net/units/n_tb");
    return MSCHED_OK;
}

template <int NT1>
int launch_ppo_nt1(const PpoArgs &a, dim3 grid, size_t smem, cudaStream_t s)
{
    auto k = a.A <= 8 ? ppo_grad_kernel<NT1, 1> : ppo_grad_kernel<NT1, 2>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    k<<<grid, 128, smem, s>>>(a);
    return 0;
}

}  // namespace

extern "C" {

int msched_ppo_workspace_bytes(const MschedPpoBatch *b, uint64_t *bytes)
{
    if (int rc = ppo_check(b)) return rc;
    if (!bytes) return fail(MSCHED_E_ARG, "null bytes");
    const int P = ppo_param_count(b->n_in, b->n_actions) + ppo_param_count(b->n_in, 1) + 4;
    *bytes = (uint64_t)b->n_sel * ppo_grid_x(b) * P * sizeof(float);
    return MSCHED_OK;
}

int msched_ppo_grad(const MschedPpoBatch *b, void *stream)
{
    if (int rc = ppo_check(b)) return rc;
    if (!b->actor_weights || !b->critic_weights || !b->x || !b->action || !b->logprob_old || !b->returns || !b->net_ids ||
        !b->unit_ids || !b->grad_actor || !b->grad_critic || !b->workspace)
        return fail(MSCHED_E_ARG, "ppo_grad: null buffer");
    const int pcA = ppo_param_count(b->n_in, b->n_actions), pcC = ppo_param_count(b->n_in, 1);
    const int P = pcA + pcC + 4;
    const int gx = ppo_grid_x(b);
    if (b->workspace_bytes < (uint64_t)b->n_sel * gx * P * sizeof(float)) return fail(MSCHED_E_ARG, "ppo_grad: workspace too small");
    PpoArgs a;
    a.actorW = b->actor_weights; a.criticW = b->critic_weights;
    a.x = b->x; a.xTbStride = b->x_tb_stride; a.xUnitStride = b->x_unit_stride;
    a.action = b->action; a.logpOld = b->logprob_old; a.ret = b->returns;
    a.netIds = b->net_ids; a.unitIds = b->unit_ids;
    a.nTb = b->n_tb; a.U = b->units; a.nSel = b->n_sel; a.m = b->units_per_net; a.nIn = b->n_in; a.A = b->n_actions;
    a.epsClip = b->eps_clip; a.entCoef = b->entropy_coef; a.valCoef = b->value_coef;
    a.partial = static_cast<float *>(b->workspace);
    int nt1 = (b->n_in + 7) / 8;  // input tiles, rounded up to an instantiated count (1, 2, 4, 8)
    nt1 = nt1 <= 2 ? nt1 : (nt1 <= 4 ? 4 : 8);
    auto fl = [](int nIn, int A) { const int Ap = (A + 3) & ~3; return (nIn * 16 + 16 + 256 + 16 + 16 * Ap + Ap + 3) & ~3; };
    size_t bufFloats = (size_t)4 * (nt1 * 8 + 48) * kPpoStride;  // four warps' buffers; the CTA sum overlays them
    if (bufFloats < (size_t)((P + 3) & ~3)) bufFloats = (size_t)((P + 3) & ~3);
    const size_t smem = sizeof(float) * ((size_t)fl(b->n_in, b->n_actions) + fl(b->n_in, 1) + bufFloats);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid(gx, b->n_sel);
    int rc = -1;
    switch (nt1) {
        case 1: rc = launch_ppo_nt1<1>(a, grid, smem, s); break;
        case 2: rc = launch_ppo_nt1<2>(a, grid, smem, s); break;
        case 3: case 4: rc = launch_ppo_nt1<4>(a, grid, smem, s); break;
        case 5: case 6: case 7: case 8: rc = launch_ppo_nt1<8>(a, grid, smem, s); break;
    }
    if (rc == -2) return fail(MSCHED_E_CUDA, "ppo_grad: shared-memory attribute rejected");
    if (rc) return fail(MSCHED_E_ARG, "ppo_grad: unsupported shape");
    CUDA_TRY(cudaGetLastError());
    const double M = (double)b->n_tb * b->units_per_net;
    ppo_reduce_kernel<<<dim3((P + 127) / 128, b->n_sel), 128, 0, s>>>(a.partial, gx, pcA, pcC, b->net_ids, (float)(1.0 / M), (float)M,
                                                                    b->grad_actor, b->grad_critic, b->stats);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, void *stream)
{
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return fail(MSCHED_E_ARG, "adam_step: bad arguments");
    if (n == 0) return MSCHED_OK;
    const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
        (float)std::sqrt(bc2), (float)eps);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

}  // extern "C"
