// msched_rollout.cu -- msched_policy_step (include/msched.h): every PPO unit of a rollout step in one launch.
#include <cstring>

#include "abi_common.h"
#include <cstdlib>

#include "policy_step_kernel.cuh"
#include "policy_step_tc_kernel.cuh"

using namespace msched;

namespace {

PolicyGroupArgs group_args(const MschedPolicyGroup &g)
{
    PolicyGroupArgs a;
    memset(&a, 0, sizeof(a));
    a.weights = g.nets.weights;
    a.nIn = g.nets.n_in; a.nActions = g.nets.n_actions; a.nNets = g.nets.n_nets;
    a.unitDiv = g.nets.unit_div > 0 ? g.nets.unit_div : 1;
    a.units = g.units; a.xOffset = g.x_offset; a.xStride = g.x_stride; a.recOffset = g.rec_offset;
    a.seed = g.seed; a.action = g.action; a.logprob = g.logprob; a.xUsed = g.x_used; a.xUsedStride = g.x_used_stride;
    a.uOverride = g.u_override; a.probs = g.probs;
    return a;
}

int g_sms = 0;

// KW = 32-bit words of an observation row (incl. the leading pad value of the acceptor rows)
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P>
int launch(PolicyStepArgs &a, cudaStream_t s)
{
    auto fn = policy_step_kernel<KW_A, AP_A, KW_O, AP_O, AP_P>;
    const size_t smem = sizeof(float) * (size_t)PolicyStepSmem<KW_A, AP_A, KW_O, AP_O, AP_P>::kWords;
    static int perSm = 0;
    if (!perSm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 128, smem) != cudaSuccess || perSm < 1) perSm = 1;
    }
    // one resident wave, split between the acceptor units and the (heavier) offer units by their multiply-adds
    const int nTiles = (a.nEnvs + 255) / 256;
    const double ca = (double)a.acc.units * (32.0 * KW_A + 256 + 16.0 * AP_A);
    const double co = (double)a.core.units * ((32.0 * KW_O + 256 + 16.0 * AP_O) + (AP_P > 0 ? (64.0 + 256 + 16.0 * AP_P) : 0.0));
    const int total = g_sms * perSm;
    int na = (int)(total * ca / (ca + co) / a.acc.units), no = (int)(total * co / (ca + co) / a.core.units);
    na = na < 1 ? 1 : (na > nTiles ? nTiles : na);
    no = no < 1 ? 1 : (no > nTiles ? nTiles : no);
    a.ctasPerAccUnit = na;
    a.ctasPerOffUnit = no;
    fn<<<a.acc.units * na + a.core.units * no, 128, smem, s>>>(a);
    return 0;
}

// the tensor-core kernel: one row per thread, SLOTS 128-environment tiles in flight per CTA, MINB CTAs per SM
// EXACT: the action counts the instantiation is built for (acceptor | core << 8 | price << 16), 0 = any up to AP_*
template <int KW_A, int AP_A, int KW_O, int AP_O, int AP_P, int SLOTS, int MINB, int EXACT = 0>
int launch_tc(PolicyStepArgs &a, cudaStream_t s)
{
    auto fn = policy_step_tc_kernel<KW_A, AP_A, KW_O, AP_O, AP_P, SLOTS, MINB, EXACT>;
    using SM = PolicyStepTcSmem<KW_A, AP_A, KW_O, AP_O, AP_P, SLOTS>;
    const size_t smem = (size_t)SM::kBytes;
    constexpr int threads = SLOTS * 128;
    static int perSm = 0;
    if (!perSm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        // resident CTAs per SM: the launch bounds hold the registers to MINB CTAs; shared memory (dynamic + static
        // barriers + 1 KB the system reserves per CTA) and the 512 tensor-memory columns are counted here (the
        // occupancy API answered 1 for these kernels where ncu reports 2 or 3 resident CTAs)
        int smemSm = 0;
        cudaDeviceGetAttribute(&smemSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        perSm = MINB;
        const int bySmem = smemSm / ((int)smem + 1024 + 256);
        if (perSm > bySmem) perSm = bySmem;
        if (perSm > 512 / (int)SM::kTmemCols) perSm = 512 / (int)SM::kTmemCols;
        if (perSm < 1) perSm = 1;
    }
    const int nTiles = (a.nEnvs + 127) / 128;
    // per-tile cost of a unit in thread instructions, from the ncu instruction counts (profiles/r02_policy_step_tc*):
    // row load + draws + emit, and per net two Tanh epilogues (175 each), the sampling epilogue, three issue / wait rounds
    constexpr double kNet8 = 350 + 130 + 36, kNet16 = 350 + 190 + 36;
    const double ca = (double)a.acc.units * (8.0 * (KW_A <= 4 ? 4 : KW_A <= 8 ? 8 : 16) + 85 + (AP_A > 8 ? kNet16 : kNet8));
    const double co = (double)a.core.units * (8.0 * (KW_O <= 4 ? 4 : KW_O <= 8 ? 8 : 16) + 85 + (AP_O > 8 ? kNet16 : kNet8) +
                                              (AP_P > 0 ? 35 + (AP_P > 8 ? kNet16 : kNet8) : 0.0));
    // CTAs per acceptor unit (na) and per offer unit (no): a slot walks ceil(tiles / (CTAs * SLOTS)) tiles of its unit, and
    // the launch lasts as long as its slowest slot -- so the split is chosen on the WHOLE-TILE counts (at 512 tiles, 18 / 31
    // CTAs per unit mean 8 and 5 tiles per slot, 16 / 32 mean 8 and 4: a fifth less for the offer units' six-layer tiles)
    const int total = g_sms * perSm;
    const double tca = ca / a.acc.units, tco = co / a.core.units;  // cost of one tile of an acceptor / offer unit
    int na = 1, no = 1;
    double best = 1e300;
    for (int x = 1; x <= nTiles && x * a.acc.units < total; ++x) {
        int y = (total - x * a.acc.units) / a.core.units;
        if (y < 1) break;
        if (y > nTiles) y = nTiles;
        const int ta = (nTiles + x * SLOTS - 1) / (x * SLOTS), to = (nTiles + y * SLOTS - 1) / (y * SLOTS);
        const double cost = ta * tca > to * tco ? ta * tca : to * tco;
        if (cost < best - 1e-9) { best = cost; na = x; no = y; }
    }
    {   // the fewest CTAs that still give these tile counts (less staging, fewer resident warps for the same work)
        const int ta = (nTiles + na * SLOTS - 1) / (na * SLOTS), to = (nTiles + no * SLOTS - 1) / (no * SLOTS);
        while (na > 1 && (nTiles + (na - 1) * SLOTS - 1) / ((na - 1) * SLOTS) == ta) --na;
        while (no > 1 && (nTiles + (no - 1) * SLOTS - 1) / ((no - 1) * SLOTS) == to) --no;
    }
    a.ctasPerAccUnit = na;
    a.ctasPerOffUnit = no;
    fn<<<a.acc.units * na + a.core.units * no, threads, smem, s>>>(a);
    return 0;
}

}  // namespace

extern "C" int msched_policy_step(const MschedPolicyStep *ps, void *stream)
{
    if (!ps || !ps->obs) return fail(MSCHED_E_ARG, "null policy step / observations");
    const MschedPolicyGroup &A = ps->acceptor, &O = ps->core, &P = ps->price;
    if (!A.nets.weights || !O.nets.weights) return fail(MSCHED_E_ARG, "acceptor and core / offer nets are required");
    const bool free = P.nets.weights != nullptr;
    if (A.nets.n_hidden != 16 || O.nets.n_hidden != 16 || (free && P.nets.n_hidden != 16))
        return fail(MSCHED_E_ARG, "msched_policy_step serves the 16-wide nets (use msched_actor_forward)");
    if (ps->n_envs < 0 || A.units < 1 || O.units < 1 || A.nets.n_nets < 1 || O.nets.n_nets < 1 || (free && (P.units != O.units || P.nets.n_nets < 1)))
        return fail(MSCHED_E_ARG, "bad n_envs/units/n_nets");
    if ((ps->env_offset & 1) || (ps->obs_stride & 1)) return fail(MSCHED_E_ARG, "env_offset and obs_stride must be even");
    const int lead = A.x_offset & 1;
    if (((A.x_offset - lead) & 1) || (A.x_stride & 1) || (O.x_offset & 1) || (O.x_stride & 1))
        return fail(MSCHED_E_ARG, "observation rows must be word aligned (msched_get_layout offsets)");
    if (ps->n_cores < 1 || O.nets.n_in != 2 * ps->n_cores + 2 || O.nets.n_actions != ps->n_cores + 1)
        return fail(MSCHED_E_ARG, "offer rows are [2*n_cores+2] with n_cores+1 actions");
    if (free && P.nets.n_in != 4) return fail(MSCHED_E_ARG, "price chooser: n_in == 4");
    if ((A.x_used && (A.x_used_stride & 1)) || (O.x_used && (O.x_used_stride & 1)) || (free && P.x_used && P.x_used_stride < 4))
        return fail(MSCHED_E_ARG, "x_used_stride must be even (whole words)");
    if (ps->n_envs == 0) return MSCHED_OK;
    PolicyStepArgs a;
    memset(&a, 0, sizeof(a));
    a.obs = ps->obs; a.obsStride = ps->obs_stride; a.nEnvs = ps->n_envs; a.nCores = ps->n_cores;
    a.actionRec = ps->action_rec; a.actionRecStride = ps->action_rec_stride; a.envOffset = ps->env_offset;
    a.step = ps->step; a.stepDev = reinterpret_cast<const unsigned long long *>(ps->step_dev);
    a.acc = group_args(A); a.core = group_args(O);
    if (free) a.price = group_args(P);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ka = A.nets.n_in, aa = A.nets.n_actions, ko = O.nets.n_in, ao = O.nets.n_actions, ap = free ? P.nets.n_actions : 0;
    int rc = -1;
    // The tensor-core kernel takes the observations as exact fp16 operands (|x| <= 511: the caller states its bound in
    // input_bound); anything else runs the fp32 SIMT kernel.  MSCHED_POLICY_STEP_IMPL=tc|simt forces one (the parity
    // tests run both).  Measured at 65,536 envs, cfg3 / cfg2 shapes: 47 / 94 us tensor cores, 78 / 225 us SIMT
    bool tc = ps->input_bound > 0 && ps->input_bound <= 511;
    if (const char *e = getenv("MSCHED_POLICY_STEP_IMPL")) {
        if (!strcmp(e, "simt")) tc = false;
        else if (!strcmp(e, "tc") && !tc) return fail(MSCHED_E_ARG, "MSCHED_POLICY_STEP_IMPL=tc needs 0 < input_bound <= 511");
    }
    if (tc) {
        // the BASELINE shapes have instantiations built for their exact action counts (config 3: 7 / 4 / 9 actions, configs
        // 2 and 4: 13 / 5), whose sampling epilogues skip the padding columns
        if (lead == 1 && ka == 15 && aa <= 8 && ko == 8 && ao <= 8 && free && ap <= 16 && (!A.x_used || A.x_used_stride >= 16) && (!O.x_used || O.x_used_stride >= 8))
            rc = (aa == 7 && ao == 4 && ap == 9) ? launch_tc<8, 8, 4, 8, 16, 4, 2, (7 | (4 << 8) | (9 << 16))>(a, s)
                                                 : launch_tc<8, 8, 4, 8, 16, 4, 2>(a, s);
        else if (lead == 1 && ka == 27 && aa <= 16 && ko == 10 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 28) && (!O.x_used || O.x_used_stride >= 10))
            rc = (aa == 13 && ao == 5) ? launch_tc<14, 16, 5, 8, 0, 4, 2, (13 | (5 << 8))>(a, s) : launch_tc<14, 16, 5, 8, 0, 4, 2>(a, s);
        else if (lead == 1 && ka == 15 && aa <= 8 && ko == 8 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 16) && (!O.x_used || O.x_used_stride >= 8))
            rc = launch_tc<8, 8, 4, 8, 0, 4, 2>(a, s);
        else if (lead == 1 && ka == 11 && aa <= 8 && ko == 8 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 12) && (!O.x_used || O.x_used_stride >= 8))
            rc = launch_tc<6, 8, 4, 8, 0, 4, 2>(a, s);
        if (rc == -2) return fail(MSCHED_E_CUDA, "msched_policy_step: shared-memory attribute rejected");
        if (rc) return fail(MSCHED_E_ARG, "msched_policy_step: no kernel for these net shapes (use msched_actor_forward per group)");
        CUDA_TRY(cudaGetLastError());
        return MSCHED_OK;
    }
    // BASELINE cfg3 (N2 C3 L3, free prices): acceptor 15 -> 7, core chooser 8 -> 4, price chooser 4 -> <= 16
    if (lead == 1 && ka == 15 && aa <= 8 && ko == 8 && ao <= 8 && free && ap <= 16 && (!A.x_used || A.x_used_stride >= 16) && (!O.x_used || O.x_used_stride >= 8))
        rc = launch<8, 8, 4, 8, 16>(a, s);
    // BASELINE cfg2 / cfg4 (N4 C4 L3, fixed prices): acceptor 27 -> 13, offer 10 -> 5
    else if (lead == 1 && ka == 27 && aa <= 16 && ko == 10 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 28) && (!O.x_used || O.x_used_stride >= 10))
        rc = launch<14, 16, 5, 8, 0>(a, s);
    // cfg3 domain with fixed prices
    else if (lead == 1 && ka == 15 && aa <= 8 && ko == 8 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 16) && (!O.x_used || O.x_used_stride >= 8))
        rc = launch<8, 8, 4, 8, 0>(a, s);
    // BASELINE cfg1 domain (N2 C3 L2): acceptor 11 -> 5, offer 8 -> 4
    else if (lead == 1 && ka == 11 && aa <= 8 && ko == 8 && ao <= 8 && !free && (!A.x_used || A.x_used_stride >= 12) && (!O.x_used || O.x_used_stride >= 8))
        rc = launch<6, 8, 4, 8, 0>(a, s);
    if (rc == -2) return fail(MSCHED_E_CUDA, "msched_policy_step: shared-memory attribute rejected");
    if (rc) return fail(MSCHED_E_ARG, "msched_policy_step: no kernel for these net shapes (use msched_actor_forward per group)");
    CUDA_TRY(cudaGetLastError());
    return MSCHED_OK;
}
