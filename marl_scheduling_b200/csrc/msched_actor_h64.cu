// msched_actor_h64.cu -- one kernel family of msched_actor_forward (include/msched.h) in its own translation unit: the build is as
// long as its slowest unit, and the actor kernels were four minutes in one.  Explicit instantiations of the launchers in
// actor_tc_wide_kernel.cuh; msched_policy.cu holds the dispatcher.
#include "abi_common.h"
#include "msched_common.cuh"
#include "policy_kernels.cuh"
#include "actor_tc_kernel.cuh"
#include "actor_tc_wide_kernel.cuh"

namespace msched {
template int launch_actor_h<64>(const ActorArgs &, const MschedMlpGroup &, dim3, int, cudaStream_t);
}  // namespace msched
