// warp_step.h -- host-side entry points of the warp-per-environment step kernel (msched_warp.cu)
#pragma once
#include <cuda_runtime.h>

#include "msched_common.cuh"

namespace msched {
// fills this translation unit's constant tables; call once per device before the first launch
cudaError_t warp_step_init();
// shared memory of one 4-environment CTA; 0 if the domain is not served (C > 64, L > 32, scratch too large)
size_t warp_step_smem_bytes(const DevParams &p, int smemOptin);
cudaError_t warp_step_prepare(size_t smem);
void launch_warp_step(const DevParams &p, size_t smem, cudaStream_t s);
void launch_observe_compact(const DevParams &p, cudaStream_t s);
int compact_obs_halfs(int C, int NL);
}  // namespace msched
