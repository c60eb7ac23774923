// msched_warp.cu -- the warp-per-environment step kernel (large domains, BASELINE config 5) and the
// compact observation kernel, in their own translation unit (include/msched.h: msched_step,
// msched_step_compact, msched_observe_compact reach them through warp_step.h).
#include "warp_step.h"
#include "warp_step_kernel.cuh"

namespace msched {

cudaError_t warp_step_init()
{
    double rcp[256];
    unsigned char odd[256];
    uint32_t magic[256];
    for (int t = 0; t < 256; ++t) {
        rcp[t] = t ? 1.0 / (double)t : 0.0;
        int o = t ? t : 1;
        while ((o & 1) == 0) o >>= 1;
        odd[t] = (unsigned char)o;
        magic[t] = t > 1 ? (uint32_t)((0x100000000ull + (unsigned long long)t - 1ull) / (unsigned long long)t) : 0u;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_rcp, rcp, sizeof(rcp));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_oddpart, odd, sizeof(odd));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_magic, magic, sizeof(magic));
}

size_t warp_step_smem_bytes(const DevParams &p, int smemOptin)
{
    if (p.C > 64 || p.L > 32 || p.NL > 4096 || p.N > 250) return 0;
    const size_t b = warp_step_smem(p.W, p.N, p.C, p.NL);
    return b + 64 <= (size_t)smemOptin ? b : 0;
}

cudaError_t warp_step_prepare(size_t smem)
{
    return cudaFuncSetAttribute(warp_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

void launch_warp_step(const DevParams &p, size_t smem, cudaStream_t s)
{
    warp_step_kernel<<<p.Bpad / 4, 128, smem, s>>>(p);
}

void launch_observe_compact(const DevParams &p, cudaStream_t s)
{
    observe_compact_kernel<<<(p.B + 3) / 4, 128, 0, s>>>(p);
}

int compact_obs_halfs(int C, int NL) { return compact_halfs(C, NL); }

}  // namespace msched
