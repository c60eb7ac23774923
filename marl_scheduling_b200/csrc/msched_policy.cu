// msched_policy.cu -- the policy side of the C-ABI (include/msched.h): actor forward on the tensor
// cores (or the fp32 SIMT kernel) and the discounted-returns kernel.  No CPU fallback.
#include <cstdlib>
#include <cstring>
#include <cmath>

#define MSCHED_ACTOR_DISPATCH_TU  // the 32- / 64-wide nets and the aggregated heads are compiled in their own units
#include "abi_common.h"
#include "msched_common.cuh"
#include "policy_kernels.cuh"
#include "actor_tc_kernel.cuh"
#include "actor_tc_wide_kernel.cuh"

using namespace msched;


namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime (no link against libcuda); null if the driver does not have it
EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
}  // namespace

extern "C" {

int msched_mlp_param_count(int n_in, int n_hidden, int n_actions)
{
    return n_hidden * n_in + n_hidden + n_hidden * n_hidden + n_hidden + n_actions * n_hidden + n_actions;
}

int msched_actor_forward(const MschedMlpGroup *nets, const MschedActorIO *io, void *stream)
{
    if (!nets || !io || !io->x || !nets->weights) return fail(MSCHED_E_ARG, "null nets/io/x/weights");
    if (io->n_envs < 0 || io->units < 1 || nets->n_nets < 1)
        return fail(MSCHED_E_ARG, "bad n_envs/units/n_nets");
    if (nets->n_in < 1 || nets->n_in > 128 || nets->n_hidden < 8 || nets->n_hidden > 64 ||
        (nets->n_hidden % 8) != 0 || nets->n_actions < 1)
        return fail(MSCHED_E_ARG, "unsupported MLP shape (in 1..128, hidden 8..64 multiple of 8)");
    if (io->gather_core) {
        if (nets->n_in != 4 || io->n_cores < 1 || io->x_stride < 2 * io->n_cores + 2)
            return fail(MSCHED_E_ARG, "gather_core needs n_in == 4 and offer observation rows of 2*n_cores+2");
    } else if (io->x_stride < nets->n_in) {
        return fail(MSCHED_E_ARG, "x_stride smaller than n_in");
    }
    if (io->action_rec && io->action_rec_stride < io->units)
        return fail(MSCHED_E_ARG, "action_rec_stride smaller than units");
    if (io->n_envs == 0) return MSCHED_OK;
    // Tensor cores (tcgen05, 3xTF32) where the contraction is wide enough to pay for the operand staging and the
    // three MMA round trips per tile: hidden width >= 32 or more than 16 actions (measured on B200, 65,536 envs:
    // 12->32->32->64 net 33.9 us vs 56.8 us SIMT; 15->16->16->7 net 41 us vs 35 us SIMT).  The 16-wide nets of a
    // whole rollout step are served by msched_policy_step; this entry point is the per-group form (aggregated
    // heads, shapes without a one-launch kernel, tests).  MSCHED_ACTOR_IMPL=tc|simt forces one.
    int impl = (nets->n_hidden >= 32 || nets->n_actions > 16) ? 0 : 1;
    if (const char *e = getenv("MSCHED_ACTOR_IMPL")) impl = !strcmp(e, "simt") ? 1 : (!strcmp(e, "tc") ? 0 : impl);
    if (impl == 1 && nets->n_hidden > 16) impl = 0;  // the SIMT kernel is built for the 16-wide nets (the wider ones always run
                                                     // on the tensor cores; their SIMT instantiations were 3.5 minutes of build)
    int rc = launch_actor_forward(*nets, *io, impl, static_cast<cudaStream_t>(stream));
    if (rc == -1) return fail(MSCHED_E_ARG, "unsupported MLP shape for the actor kernel");
    if (rc == -2) return fail(MSCHED_E_CUDA, "actor kernel: shared-memory attribute rejected");
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_returns(const float *rewards, int T, int M, double gamma, int normalise, float *out, void *stream)
{
    if (T < 1 || M < 0) return fail(MSCHED_E_ARG, "bad T/M");
    if (normalise && T < 2) return fail(MSCHED_E_ARG, "normalisation needs T >= 2");
    if (M == 0) return MSCHED_OK;  // an empty buffer has no address to check
    if (!rewards || !out) return fail(MSCHED_E_ARG, "null rewards/out");
    // One 2-D tensor-map copy per tile and direction when the buffer qualifies (T <= 256 rows in the box, the [T][128]
    // tile in shared memory, 16-byte aligned rows); MSCHED_RETURNS_TMAP=0 keeps the row-by-row copies
    if ((M & 3) == 0 && T <= 256 && (size_t)T * 128 * sizeof(float) <= 200 * 1024 && (reinterpret_cast<uintptr_t>(rewards) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && !(getenv("MSCHED_RETURNS_TMAP") && atoi(getenv("MSCHED_RETURNS_TMAP")) == 0)) {
        if (EncodeTiledFn enc = tensor_map_encoder()) {
            // tile width: 64 columns, 128 for short buffers (more tiles resident per SM: the loads, scans and stores of
            // different tiles overlap, and with one copy per tile a narrower box costs the copy unit nothing).  Measured,
            // 128 / 64 / 32 columns: T = 200 x 393,216 units 203 / 188 / 184 us, T = 64 x 1,048,576 97 / 92 / 101 us,
            // T = 16 x 393,216 15.0 / 15.4 / 15.6 us
            int W = T <= 32 ? 128 : 64;
            if (const char *e = getenv("MSCHED_RETURNS_W")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) W = v; }
            CUtensorMap tin, tout;
            const cuuint64_t gdim[2] = {(cuuint64_t)M, (cuuint64_t)T}, gstr[1] = {(cuuint64_t)M * sizeof(float)};
            const cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)T}, estr[2] = {1u, 1u};
            const CUresult r1 = enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(rewards), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const CUresult r2 = enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) {
                static bool attrT = false;
                if (!attrT) {
                    CUDA_TRY(cudaFuncSetAttribute(returns_tmap_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                    CUDA_TRY(cudaFuncSetAttribute(returns_tmap_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                    CUDA_TRY(cudaFuncSetAttribute(returns_tmap_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                    attrT = true;
                }
                const size_t tb = (size_t)T * W * sizeof(float);
                cudaStream_t s = static_cast<cudaStream_t>(stream);
                if (W == 128) returns_tmap_kernel<128><<<(M + 127) / 128, 128, tb, s>>>(tin, tout, T, M, gamma, normalise);
                else if (W == 64) returns_tmap_kernel<64><<<(M + 63) / 64, 64, tb, s>>>(tin, tout, T, M, gamma, normalise);
                else returns_tmap_kernel<32><<<(M + 31) / 32, 32, tb, s>>>(tin, tout, T, M, gamma, normalise);
                CUDA_TRY(cudaGetLastError());
                return MSCHED_OK;
            }
        }
    }
    // TMA-tiled kernel (every reward read once) when the rows are 16-byte aligned and a [T][W] tile fits in shared
    // memory -- W = 128 columns up to T = 400, 64 up to 800, 32 up to 1,600 (narrower tiles are not faster: 261 / 263 /
    // 284 us at T = 200, the float64 chain of a column is the bound) -- the streaming kernel otherwise
    if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(rewards) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        int W = 0;
        for (int w = 128; w >= 32 && !W; w >>= 1)
            if ((size_t)T * w * sizeof(float) <= 200 * 1024) W = w;
        if (const char *e = getenv("MSCHED_RETURNS_W")) { const int v = atoi(e); if ((v == 32 || v == 64 || v == 128 || v == 256) && (size_t)T * v * sizeof(float) <= 200 * 1024) W = v; }
        if (W) {
            const size_t tileBytes = (size_t)T * W * sizeof(float);
            static bool attr = false;
            if (!attr) {
                CUDA_TRY(cudaFuncSetAttribute(returns_tile_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                CUDA_TRY(cudaFuncSetAttribute(returns_tile_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                CUDA_TRY(cudaFuncSetAttribute(returns_tile_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                CUDA_TRY(cudaFuncSetAttribute(returns_tile_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
                attr = true;
            }
            cudaStream_t s = static_cast<cudaStream_t>(stream);
            if (W == 256) returns_tile_kernel<256><<<(M + 255) / 256, 256, tileBytes, s>>>(rewards, T, M, gamma, normalise, out);
            else if (W == 128) returns_tile_kernel<128><<<(M + 127) / 128, 128, tileBytes, s>>>(rewards, T, M, gamma, normalise, out);
            else if (W == 64) returns_tile_kernel<64><<<(M + 63) / 64, 64, tileBytes, s>>>(rewards, T, M, gamma, normalise, out);
            else returns_tile_kernel<32><<<(M + 31) / 32, 32, tileBytes, s>>>(rewards, T, M, gamma, normalise, out);
            CUDA_TRY(cudaGetLastError());
            return MSCHED_OK;
        }
    }
    returns_kernel<<<(M + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(rewards, T, M, gamma,
                                                                                   normalise, out);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}

int msched_dqn_param_count(int n_in, int n_actions) { return 16 * n_in + 16 + n_actions * 16 + n_actions; }

int msched_dqn_select(const MschedMlpGroup *nets, const MschedActorIO *io, float epsilon, float *q_out, void *stream)
{
    if (!nets || !io || !io->x || !nets->weights) return fail(MSCHED_E_ARG, "null nets/io/x/weights");
    if (nets->n_hidden != 16) return fail(MSCHED_E_ARG, "the DQN nets have 16 hidden neurons (src/DQNmodules.py:41-46)");
    if (io->n_envs < 0 || io->units < 1 || nets->n_nets < 1 || nets->n_in < 1 || nets->n_in > 512 ||
        nets->n_actions < 1 || nets->n_actions > 1024 || io->x_stride < nets->n_in)
        return fail(MSCHED_E_ARG, "bad n_envs/units/n_nets/n_in/n_actions/x_stride");
    if (io->gather_core) return fail(MSCHED_E_ARG, "gather_core is a PPO free-price feature");
    if (io->action_rec && io->action_rec_stride < io->units) return fail(MSCHED_E_ARG, "action_rec_stride smaller than units");
    if (io->n_envs == 0) return MSCHED_OK;
    QArgs q;
    memset(&q, 0, sizeof(q));
    ActorArgs &a = q.a;
    a.weights = nets->weights; a.x = io->x;
    a.envStride = io->env_stride ? io->env_stride : (long long)io->x_stride * io->units; a.unitStride = io->x_stride;
    a.nIn = nets->n_in; a.nHidden = 16; a.nActions = nets->n_actions; a.nNets = nets->n_nets;
    a.unitDiv = nets->unit_div > 0 ? nets->unit_div : 1;
    a.units = io->units; a.nEnvs = io->n_envs;
    a.seed = io->seed; a.step = io->step; a.rowOffset = io->row_offset; a.uOverride = io->u_override;
    a.action = io->action; a.actionRec = io->action_rec; a.actionRecStride = io->action_rec_stride;
    a.stepDev = reinterpret_cast<const unsigned long long *>(io->step_dev);
    q.epsilon = epsilon;
    q.qOut = q_out;
    const size_t smem = sizeof(float) * ((size_t)16 * a.nIn + 16 + (size_t)a.nActions * 16 + a.nActions);
    int tiles = (a.nEnvs + 127) / 128;
    int gx = (148 * 8 + io->units - 1) / io->units;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(dqn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dqn_select_kernel<<<dim3(gx, io->units), 128, smem, static_cast<cudaStream_t>(stream)>>>(q);
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}


}  // extern "C"
