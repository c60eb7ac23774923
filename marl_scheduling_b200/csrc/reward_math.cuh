// reward_math.cuh -- the exact integer / float64 arithmetic of the reward functions, shared by every step kernel.
// Each translation unit that includes it owns a copy of the constant tables and must fill them once
// (msched_abi.cu: msched_create; msched_warp.cu: warp_step_init).
#pragma once
#include "msched_common.cuh"

namespace msched {

// 1/t and the odd part of t for t in 0..255 (job lengths are 1..255), filled at library load
__constant__ double c_rcp[256];
__constant__ unsigned char c_oddpart[256];

// tradedReward = round(offeredReward / necessaryTime * timeMeasure), src/Reward.py:70-71: float64
// divide, float64 multiply, Python round() = round-half-even (Q7).  Fast path: the exact rational
// price*dt/time is at least 1/(2*time) away from a rounding boundary unless it is an exact tie, and
// the float64 evaluation is within 2^-20 of it, so the nearest integer is the answer.  Exact ties
// are decided like the float64 code does when price/time is exactly representable (tie -> even);
// everything else takes the literal float64 path.
static __device__ __noinline__ int traded_reward_slow(int price, int time, int dt)
{
    const double ratio = __ddiv_rn((double)price, (double)time);
    return (int)rint(__dmul_rn(ratio, (double)dt));
}
__device__ __forceinline__ int traded_reward(int price, int time, int dt)
{
    if ((unsigned)dt < 32768u) {
        const int num = price * dt;  // |num| < 2^30
        const int m = __double2int_rn(__dmul_rn((double)num, c_rcp[time]));
        const int d2 = 2 * (num - m * time);
        const int ad = d2 < 0 ? -d2 : d2;
        if (ad < time) return m;
        if (ad == time && (price % (int)c_oddpart[time]) == 0) {
            const int lo = d2 > 0 ? m : m - 1;  // exact value is lo + 0.5
            return (lo & 1) ? lo + 1 : lo;
        }
    }
    return traded_reward_slow(price, time, dt);
}

}  // namespace msched
