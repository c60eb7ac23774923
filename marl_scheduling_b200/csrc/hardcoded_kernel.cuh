// hardcoded_kernel.cuh -- the reference's heuristic agents (BASELINE config 1), batched.
//
//   src/HardcodedModules.py:5-13    calculateRewardRatio (-1 for the -1 / -2 paddings)
//   src/HardcodedModules.py:16-45   HardcodedAcceptor.selectAction
//   src/HardcodedModules.py:81-109  HardcodedOfferer.selectAction
//   src/Agent.py:622-641            DividedHardcodedAgent.getActions
//   src/SchedulingEnvironment.py:150-172  getActionForAllAgents
//
// One thread per (environment, unit): units 0 .. N*C-1 are the acceptors (agent, core), units N*C ..
// N*C+N*L-1 the offerers (agent, slot).  Each reads its row of the dense observation record -- the
// same information the reference's policies are handed -- and writes its action into the env's
// action record.  Ratios are compared exactly by cross-multiplication (all denominators are positive
// job lengths; float division in the reference gives equal quotients iff the rationals are equal and
// orders them correctly for int16 operands).  The reference breaks ties with random.sample (global
// Mersenne Twister, consumed even for a single candidate); here candidate floor(u * #candidates) in
// index order is taken, u from Philox (counter = global env, round, stream 3, unit) or from u_override.
#pragma once
#include "msched_common.cuh"

namespace msched {

constexpr uint32_t kStreamHardcoded = 3;

struct HardcodedArgs {
    const int16_t *obs;
    int16_t *action;         // the action record (acceptor idx and offer core fields are written)
    const float *uOverride;  // [B][N*C + N*L] or null
    int32_t *ncand;          // [B][N*C + N*L] number of tie candidates (0 = no choice made), or null
    int oAcc, oOff, accRow, offRow;  // observation record offsets / row strides (int16 elements)
    int randomTies;
};

// ratio of calculateRewardRatio as an exact fraction num/den, den > 0
__device__ __forceinline__ void hc_ratio(int a, int b, int &num, int &den)
{
    if (a == -1 || b == -1 || a == -2 || b == -2) { num = -1; den = 1; }
    else { num = a; den = b; }
}

// the action of one unit (acceptors 0 .. N*C-1, offerers N*C ..) from the environment's dense observation record `ob`
// (global memory or the shared-memory tile of the fused kernel); nc = number of tie candidates (0: no choice made)
__device__ __forceinline__ int hc_unit_action(const DevParams &p, const int16_t *ob, int oAcc, int oOff, int accRow, int offRow, int unit,
                                              float u, int &nc)
{
    nc = 0;
    int act;
    if (unit < p.N * p.C) {
        // ---- HardcodedAcceptor: accept the best offeredReward/necessaryTime if it beats the own job's
        // priority/remainingLength, else reject (index NL); not the owner -> reject ----
        const int16_t *row = ob + oAcc + (size_t)unit * accRow;
        act = p.NL;
        if (row[0] != 0) {
            int on, od;
            hc_ratio(row[1], row[2], on, od);
            int bn = -1, bd = 1;  // max over the offers; paddings rate -1
            for (int k = 0; k < p.NL; ++k) {
                int n, d;
                hc_ratio(row[3 + 2 * k], row[4 + 2 * k], n, d);
                if (d <= 0) { n = -1; d = 1; }  // a non-positive time cannot occur (job lengths >= 1)
                const int lhs = n * bd, rhs = bn * d;
                if (lhs > rhs) { bn = n; bd = d; nc = 1; }
                else if (lhs == rhs) ++nc;
            }
            if (od <= 0) { on = -1; od = 1; }
            if ((long long)bn * od > (long long)on * bd) {
                int pick = (int)(u * (float)nc);
                pick = pick >= nc ? nc - 1 : pick;
                int t = 0;
                for (int k = 0; k < p.NL; ++k) {
                    int n, d;
                    hc_ratio(row[3 + 2 * k], row[4 + 2 * k], n, d);
                    if (d <= 0) { n = -1; d = 1; }
                    if (n * bd == bn * d) {
                        if (t == pick) { act = k; break; }
                        ++t;
                    }
                }
            } else {
                nc = 0;
            }
        }
    } else {
        // ---- HardcodedOfferer: offer to a (random) core with the LOWEST priority/remainingLength, an
        // idle core rating -1; never abstains ----
        const int s = unit - p.N * p.C;
        const int16_t *row = ob + oOff + (size_t)s * offRow;
        int bn = 0, bd = 0;
        for (int j = 0; j < p.C; ++j) {
            int n, d;
            hc_ratio(row[2 * j], row[2 * j + 1], n, d);
            if (d <= 0) { n = -1; d = 1; }
            if (bd == 0 || n * bd < bn * d) { bn = n; bd = d; nc = 1; }
            else if (n * bd == bn * d) ++nc;
        }
        int pick = (int)(u * (float)nc);
        pick = pick >= nc ? nc - 1 : pick;
        act = 0;
        int t = 0;
        for (int j = 0; j < p.C; ++j) {
            int n, d;
            hc_ratio(row[2 * j], row[2 * j + 1], n, d);
            if (d <= 0) { n = -1; d = 1; }
            if (n * bd == bn * d) {
                if (t == pick) { act = j; break; }
                ++t;
            }
        }
    }
    return act;
}

// the unit's tie draw: word unit % 4 of Philox call unit / 4 of the hard-coded-agents stream, at the round the
// observations belong to
__device__ __forceinline__ float hc_unit_draw(const DevParams &p, int round, int env, int unit)
{
    uint32_t x[4];
    env_draw_at(p, round, env, kStreamHardcoded, (uint32_t)(unit >> 2), 0u, x);
    const uint32_t xv = (unit & 3) == 0 ? x[0] : (unit & 3) == 1 ? x[1] : (unit & 3) == 2 ? x[2] : x[3];
    return (float)(xv >> 8) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(128) hardcoded_policy_kernel(const __grid_constant__ DevParams p, const HardcodedArgs a)
{
    const int U = p.N * p.C + p.NL;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.B * U) return;
    const int env = (int)(gid / U), unit = (int)(gid % U);
    const int16_t *ob = a.obs + (size_t)env * p.OH;
    float u = 0.f;
    if (a.uOverride) u = a.uOverride[gid];
    else if (a.randomTies) u = hc_unit_draw(p, cur_round(p), env, unit);
    int nc;
    const int act = hc_unit_action(p, ob, a.oAcc, a.oOff, a.accRow, a.offRow, unit, u, nc);
    if (unit < p.N * p.C) a.action[(size_t)env * p.AH + p.aAcc + unit] = (int16_t)act;
    else a.action[(size_t)env * p.AH + p.aOffc + unit - p.N * p.C] = (int16_t)act;
    if (a.ncand) a.ncand[gid] = nc;
}

}  // namespace msched
