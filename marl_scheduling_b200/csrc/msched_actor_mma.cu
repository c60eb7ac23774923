// msched_actor_mma.cu -- the warp-level tensor-core actor kernel for the 16-wide nets, its own translation unit
#include "abi_common.h"
#include "msched_common.cuh"
#include "actor_mma_kernel.cuh"

namespace msched {

namespace {

template <int NT1>
int launch_actor_mma(const ActorArgs &a, dim3 grid, cudaStream_t s)
{
    auto k = a.nActions <= 8 ? actor_forward_mma<NT1, 1> : actor_forward_mma<NT1, 2>;
    static int nSm = 0, perSm = 0;
    static const void *cached = nullptr;
    if (!nSm || cached != reinterpret_cast<const void *>(k)) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, 128, 0) != cudaSuccess || perSm < 1) perSm = 1;
        cached = reinterpret_cast<const void *>(k);
    }
    int gx = (nSm * perSm) / (int)grid.y;  // persistent: the weight fragments are loaded once per warp
    if (gx > (int)grid.x) gx = (int)grid.x;
    if (gx < 1) gx = 1;
    k<<<dim3(gx, grid.y), 128, 0, s>>>(a);
    return 0;
}

}  // namespace

int launch_actor_mma_any(const ActorArgs &a, dim3 grid, cudaStream_t s)
{
    const int nt1 = (a.nIn + 7) / 8;
    if (nt1 == 1) return launch_actor_mma<1>(a, grid, s);
    if (nt1 == 2) return launch_actor_mma<2>(a, grid, s);
    if (nt1 <= 4) return launch_actor_mma<4>(a, grid, s);
    return -1;
}

}  // namespace msched
