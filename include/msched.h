/*
 * msched.h -- C-ABI of the B200-native batched marl-scheduling rollout path.
 *
 * The reference (lr40/marl-scheduling) has no FFI layer: the path sits behind plain Python
 * classes (SchedulingEnv.step/reset, World, Auctioneer, the PPO units).  This header is the
 * boundary a maintainer would bind with ctypes from those classes; every entry point names
 * the reference interface it replaces (file:line relative to the reference repo).  The
 * Python binding shipped here is marl_scheduling_b200/_lib.py; INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are CALLER-OWNED DEVICE buffers unless
 *     the name ends in _host; nothing is allocated after msched_create;
 *   - every launch takes a cudaStream_t (passed as void*), is asynchronous and returns an int
 *     status: 0 ok, <0 MSCHED_E_* (argument / shape / CUDA error, see msched_last_error);
 *   - per-environment runtime faults are reported through the sticky uint32 flags word of the
 *     result record (MSCHED_FLAG_*), never by aborting;
 *   - one process per GPU, one handle per env shard; envs never communicate.
 *
 * Record streams (all env-major, one fixed-size record per environment, sizes from
 * msched_get_layout; buffers must hold msched_padded_envs(B) records):
 *   state   uint32[state_words]   opaque persistent state (cores, job slots, pending offers)
 *   action  int16 [action_halfs]  acceptor idx [N][C] | offer core [N][L] | offer price [N][L]
 *                                 | auctioneer idx [C] | spawn kind [N][newJobs]
 *   result  uint32[result_words]  offer reward f32 [N][RL] | price reward f32 [N][RL]
 *                                 | acceptor reward i32 [N][RC] | auctioneer reward i32 [C]
 *                                 | agent reward i32 [N] | quality_sum f64 (lo,hi words)
 *                                 | counts (quality_cnt | n_accepted<<8 | n_terminated<<16
 *                                 | done<<24) | flags | auctioneer idx used, 2 x i16 per word
 *   obs     int16 [obs_halfs]     dense reference-layout observations (Appendix B of SURVEY.md),
 *                                 rows padded to aligned 32-bit (value, value) pairs
 *   ids     int16 [ids_halfs]     offer-ID tables [N][C][NL] | [C][NL] (drop-in API / parity only)
 *   chain   uint32[C][chainCap][2] liability chains (src/world.py:238,285-289), touched
 *                                 only on acceptance / termination
 */
#ifndef MSCHED_H
#define MSCHED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSCHED_ABI_VERSION 2
#define MSCHED_MAX_KINDS 16
#define MSCHED_TILE_ENVS 128 /* record buffers are padded to a multiple of this */

/* status codes */
#define MSCHED_OK 0
#define MSCHED_E_ARG (-1)      /* bad argument / unsupported shape */
#define MSCHED_E_CUDA (-2)     /* CUDA runtime error, see msched_last_error */
#define MSCHED_E_NODEVICE (-3) /* no CUDA device: there is NO CPU fallback */
#define MSCHED_E_STATE (-4)    /* state / chain buffer not bound */

/* per-env fault flags (sticky, result record + state record) */
#define MSCHED_FLAG_CHAIN_OVERFLOW 1u  /* liability chain longer than chainCap (entry dropped) */
#define MSCHED_FLAG_COLLECTION_FULL 2u /* reference: raise in JobCollection.insertJob, src/world.py:133 */
#define MSCHED_FLAG_ACTION_RANGE 4u    /* reference: assert in src/world.py:389,404 */
#define MSCHED_FLAG_SPAWN_RANGE 8u     /* reference: UnboundLocalError in src/Agent.py:52-57 (Q13) */
#define MSCHED_FLAG_COMPACT_RANGE 16u  /* compact result record only: an integer reward did not fit int16 (saturated) */

/* reward variants: src/Reward.py:146-212, :6-89 (commercial / non-commercial), :92-143 */
enum {
    MSCHED_REWARD_DIVIDED_FIXED = 0,
    MSCHED_REWARD_DIVIDED_FREE_COMMERCIAL = 1,
    MSCHED_REWARD_DIVIDED_FREE_NONCOMMERCIAL = 2,
    MSCHED_REWARD_AGGREGATED_FIXED = 3
};

/* who supplies the auctioneer's acceptor action (src/Auctioneer.py:95-102) */
enum {
    MSCHED_AUCTION_EXTERNAL = 0,     /* action record carries auctioneer idx [C] (parity mode) */
    MSCHED_AUCTION_FIRST_MAX = 1,    /* in-kernel HardcodedAuctioneerAcceptor, first arg-max */
    MSCHED_AUCTION_RANDOM_MAX = 2    /* in-kernel, uniformly random arg-max (Philox) */
};

/* where spawn draws come from (src/Agent.py:50-70) */
enum {
    MSCHED_SPAWN_PHILOX = 0, /* device Philox4x32-10, counter (global env, round, agent, k) */
    MSCHED_SPAWN_KINDS = 1,  /* action record carries the job kind per (agent, k) */
    MSCHED_SPAWN_U64 = 2     /* float64 draws u in [0,1) per (agent, k): recorded random.random() */
};

/* World(params) + SchedulingEnv(world, params): src/world.py:210-254,
 * src/SchedulingEnvironment.py:22-30,253-259 */
typedef struct MschedConfig {
    int32_t abi_version;   /* MSCHED_ABI_VERSION */
    int32_t B;             /* environments in this shard */
    int32_t N, C, L, J;    /* numberOfAgents, numberOfCores, collectionLength, #job kinds */
    int32_t newJobsPerRound, rewardMultiplier, episodeLength;
    int32_t freePrices;    /* 0/1 */
    int32_t rewardVariant; /* MSCHED_REWARD_* */
    int32_t chainCapacity; /* liability-chain entries kept per core (default 32) */
    int32_t auctionMode;   /* MSCHED_AUCTION_* */
    int32_t spawnMode;     /* MSCHED_SPAWN_* */
    int32_t prio[MSCHED_MAX_KINDS];     /* possibleJobPriorities */
    int32_t len[MSCHED_MAX_KINDS];      /* possibleJobLengths (1..255) */
    int32_t fixPrice[MSCHED_MAX_KINDS]; /* fixPricesList (fixed prices only) */
    double cumProb[MSCHED_MAX_KINDS];   /* World.accProbabilities: float64 prefix sums */
    double netZeroOfferReward;          /* src/Reward.py:29-33 */
    uint64_t seed;                      /* Philox key */
    int64_t envOffset;                  /* global index of env 0 of this shard (multi-GPU) */
} MschedConfig;

/* record geometry; element offsets inside each record (-1 = absent) */
typedef struct MschedLayout {
    int32_t padded_envs;
    int32_t state_words;  /* uint32 per env */
    int32_t action_halfs; /* int16 per env (even; action_halfs/2 is odd) */
    int32_t result_words; /* uint32 per env (odd) */
    int32_t obs_halfs;    /* int16 per env (even; obs_halfs/2 is odd) */
    int32_t ids_halfs;    /* int16 per env = (N*C + C) * N*L */
    int32_t chain_words;  /* uint32 per env = C*chainCapacity*2 */
    /* action record, int16 element offsets */
    int32_t a_acceptor, a_offer_core, a_offer_price, a_auctioneer, a_spawn_kind;
    /* result record, word offsets */
    int32_t r_offer, r_price, r_acceptor, r_auctioneer, r_agent, r_quality, r_counts, r_flags,
        r_auctioneer_idx;
    int32_t RL, RC; /* reward row lengths: L,C (divided) or 1,1 (aggregated) */
    /* obs record, int16 element offsets of the first logical element of each block and the row
     * strides: acceptor [N][C] rows of 3+2NL values, stride o_acc_row (o_acceptor is odd: the
     * rows carry one leading pad so that (price,time) pairs are aligned words); auctioneer [C]
     * rows, same stride; offer [N][L] rows of 2C+2 values, stride o_off_row */
    int32_t o_acceptor, o_offer, o_auctioneer, o_acc_row, o_off_row;
    /* compact observation record (msched_observe_compact / msched_step_compact), int16 per env:
     * core [C][4] ownerID (0 = auctioneer), priority, remainingLength, jobKind (-1 = idle) at c_core;
     * slot [N*L][2] priority, remainingLength (-1 = empty) at c_slot;
     * offer [N*L][2] coreID (0 = no pending offer), offeredReward at c_offer -- the recipient of an offer is the
     * owner of its core and its necessaryTime the slot's remainingLength (src/world.py:406-478) */
    int32_t cobs_halfs, c_core, c_slot, c_offer;
} MschedLayout;

int msched_abi_version(void);
const char *msched_last_error(void);
int msched_padded_envs(int B);
int msched_get_layout(const MschedConfig *cfg, MschedLayout *out);

/* World.__init__ / SchedulingEnv.__init__ (src/world.py:210-254) */
int msched_create(const MschedConfig *cfg, int device, void **handle);
int msched_destroy(void *handle);

/* which kernels this handle launches (diagnostics; bench.py reports it) */
typedef struct MschedInfo {
    int32_t step_impl;        /* 0 one lane per env (any domain), 1 cooperative G lanes per env,
                                 2 register-resident compile-time-domain kernel, 3 one warp per env (large domains) */
    int32_t fuses_observations; /* msched_step_observe is ONE launch */
    int32_t envs_per_cta, threads_per_cta;
    int32_t smem_bytes_per_cta; /* dynamic shared memory of the step launch (with observations if fused) */
    int32_t reserved[3];
} MschedInfo;
int msched_get_info(void *handle, MschedInfo *out);

/* diagnostics: when timeline_dev != NULL the compile-time-domain step kernel records, per CTA,
 * 8 x uint64: SM id, %globaltimer at start, clock64 at start / after the pre-work / after the tile
 * arrived / before the bulk stores / after they have read shared memory, %globaltimer at the end.
 * Buffer: padded_envs/32 * 8 uint64.  NULL switches it off (default). */
int msched_debug_timeline(void *handle, uint64_t *timeline_dev);

/* bind the caller-owned persistent buffers (state: padded_envs*state_words uint32,
 * chain: padded_envs*chain_words uint32) */
int msched_bind_state(void *handle, void *state_dev, void *chain_dev);

/* ---- episode aggregates (src/trainPPO.py:172-227: what the train scripts accumulate per step and
 * average per episode) ----
 * msched_bind_stats: caller-owned int32 [padded_envs][J][4] (zero it at episode start); every step
 *   launch then adds, per env and job kind, [sum of the accepted offers' prices, #accepted offers,
 *   sum of (dwell time - 1) of the terminated jobs, #terminated jobs] -- the raw material of the
 *   per-kind mean price and mean normalised dwell time ((dwell-1)/length, src/world.py:350-357).
 *   NULL switches it off (default).
 * msched_stats_sums: adds the column sums over the environments to out_dev int64 [J][4].
 * msched_result_sums: adds the sums over the environments of one step's result records to out_dev
 *   float64 [result_words + 6]: out[k] = sum of record word k (float fields as float, integer fields
 *   as integer, out[r_quality] = sum of quality_sum); tail: +0 sum quality_cnt, +1 sum n_accepted,
 *   +2 sum n_terminated, +3 sum done, +4 sum of the per-env mean acception quality over envs with
 *   quality_cnt > 0, +5 number of such envs. */
int msched_bind_stats(void *handle, int32_t *stats_dev);
int msched_stats_sums(void *handle, int64_t *out_dev, void *stream);
int msched_result_sums(void *handle, const uint32_t *result_dev, double *out_dev, void *stream);

/* fresh worlds: all cores auctioneer-owned, collections empty, round 0, jobIDs from 1
 * (World.__init__; note SchedulingEnv.reset itself resets nothing, src/SchedulingEnvironment.py:85-109) */
int msched_reset(void *handle, void *stream);
int msched_get_round(void *handle, int64_t *round);
int msched_set_round(void *handle, int64_t round);
/* CUDA-graph support: with device_side != 0 world.round lives in a device counter that every step
 * launch reads and a one-thread kernel advances, so a captured step can be replayed (the kernel
 * arguments no longer change from step to step).  msched_get_round then synchronises the device. */
int msched_set_round_mode(void *handle, int device_side, void *stream);

/* One SchedulingEnv.step for every env (src/SchedulingEnvironment.py:32-83 =
 * World.step1 src/world.py:295-334 + acception quality :174-192 + Reward.py + done),
 * i.e. acceptance application, auction + core allocation, job progress/completion, offer
 * creation, spawn refill, rewards.  spawn_u_dev: float64 [B][N][newJobs] for MSCHED_SPAWN_U64,
 * else NULL. */
int msched_step(void *handle, const int16_t *action_dev, const double *spawn_u_dev,
                uint32_t *result_dev, void *stream);

/* SchedulingEnv.step INCLUDING the observations it returns (src/SchedulingEnvironment.py:32-83:
 * world.step1, then every agent's gatherObservations and the auctioneer's observation of the NEW
 * state, src/SchedulingEnvironment.py:44-58): msched_step followed by msched_observe_dense, fused
 * into ONE launch when the domain has a compile-time kernel and the observation tile fits in
 * shared memory next to the state tile (two launches otherwise; same results either way). */
int msched_step_observe(void *handle, const int16_t *action_dev, const double *spawn_u_dev,
                        uint32_t *result_dev, int16_t *obs_dev, void *stream);

/* n_steps consecutive SchedulingEnv.steps in ONE launch for scripted / pre-drawn actions (rollouts of the hard-coded
 * agents' action traces, random-policy warm-up, evaluation): step t reads the action record action_dev + t *
 * padded_envs * action_halfs and writes the result record result_dev + t * padded_envs * result_words.  A CTA keeps
 * its 32 environments for all steps: no launch gap, ramp or tail between the steps of a dependent chain, and --
 * without obs_every -- the state tile stays in shared memory between steps (only the last step's observations are
 * written to obs_dev [padded_envs][obs_halfs], which may be NULL).  obs_every != 0: every step writes its
 * observations to obs_dev + t * padded_envs * obs_halfs.  Same results, bit for bit, as n_steps calls of
 * msched_step_observe; the round advances by n_steps.  Domains with a multi-step instantiation of the fused kernel
 * only (the BASELINE configurations; MSCHED_E_ARG otherwise), device Philox or recorded-kind spawn. */
int msched_step_multi(void *handle, const int16_t *action_dev, int n_steps, uint32_t *result_dev, int16_t *obs_dev,
                      int obs_every, void *stream);

/* same call with HOST buffers: the action records go in, the result records come out, then stream
 * synchronise.  PINNED buffers (cudaHostAlloc / torch pin_memory) of an unpadded batch (B a multiple of 128)
 * on a domain with a fused kernel are read and written by the kernel's bulk copies directly (zero-copy: one
 * launch, the PCIe traffic of the tiles pipelines across the resident CTAs).  Otherwise: H2D of the action
 * records, the step, D2H of the result records, the batch cut into chunks that alternate between two internal
 * streams so the copies of one chunk overlap the kernel of another.  obs_dev (DEVICE,
 * optional): the dense observations of the new state stay on the device for the policy kernels
 * (written by the same launch when the domain fuses them).  Staging buffers are owned by the handle. */
int msched_step_host(void *handle, const int16_t *action_host, uint32_t *result_host,
                     int16_t *obs_dev, void *stream);

/* msched_step_host with a COMPACT result record: what SchedulingEnv.step returns besides the observations
 * (src/SchedulingEnvironment.py:60-83: the four reward arrays of src/Reward.py, the acception quality, done) in
 * half the bytes, for callers on the far side of PCIe.  Per env `words` uint32: int16-sized planes first (offsets in
 * 16-bit units) -- offer reward and price reward as IEEE half (prio1, prio1 - price and netZeroOfferReward: exact
 * for |value| <= 2048 and a half-representable netZeroOfferReward, which msched_get_compact_result_layout checks),
 * acceptor / auctioneer / agent reward as int16 (saturated with MSCHED_FLAG_COMPACT_RANGE if one does not fit) --
 * then two words: quality_sum as float32, and counts (bits 0..24, as in the full record) with the sticky flags in
 * bits 25..31 (c_flags == c_counts).  BASELINE config 3: 56 bytes instead of 116, and a 32-env tile of 1,792 bytes
 * = 7 x 256, so every tile written over PCIe starts on a 256-byte boundary.  Needs a domain with a fused kernel, PINNED host buffers and B a multiple of 128 (else
 * MSCHED_E_ARG: use msched_step_host); the kernel's bulk copies read the actions from and write the compact
 * records to host memory directly. */
typedef struct MschedCompactResultLayout {
    int32_t words;                                             /* uint32 per env */
    int32_t c_offer, c_price, c_acceptor, c_auctioneer, c_agent; /* 16-bit element offsets (-1 = absent) */
    int32_t c_quality, c_counts, c_flags;                      /* word offsets; flags = bits 25..31 of the counts word */
} MschedCompactResultLayout;
int msched_get_compact_result_layout(const MschedConfig *cfg, MschedCompactResultLayout *out);
int msched_step_host_compact(void *handle, const int16_t *action_host, uint32_t *cresult_host, int16_t *obs_dev,
                             void *stream);

/* Agent.gatherObservations + gatherDividedAuctioneerObservation (src/Agent.py:148-300,
 * src/Auctioneer.py:20-77): dense reference-layout observations; ids_dev (optional, may be
 * NULL) receives the offer-ID tables env.correspondingOfferIDs /
 * auctioneer_correspondingOfferIDs (src/SchedulingEnvironment.py:26-29, B*ids_halfs int16) */
int msched_observe_dense(void *handle, int16_t *obs_dev, int16_t *ids_dev, void *stream);

/* Compact observations: the information of Agent.gatherObservations / gatherDividedAuctioneerObservation
 * (src/Agent.py:148-300, src/Auctioneer.py:20-77) without the per-(agent, core) replication of the dense rows --
 * cores, job slots and pending offers once each (layout: MschedLayout.cobs_halfs / c_core / c_slot / c_offer).
 * The dense record of BASELINE config 5 (N32 C64 L8) is 2.2 MB per environment, the compact one 2.5 KB; a policy
 * kernel builds an acceptor row from (core j, the offers whose coreID is j+1 in slot order) and an offer row from
 * (all cores, slot q).  cobs_dev: padded_envs * cobs_halfs int16. */
int msched_observe_compact(void *handle, int16_t *cobs_dev, void *stream);

/* msched_step followed by msched_observe_compact; ONE launch on the warp-per-environment kernel (large domains) */
int msched_step_compact(void *handle, const int16_t *action_dev, const double *spawn_u_dev, uint32_t *result_dev,
                        int16_t *cobs_dev, void *stream);

/* Auctioneer.getAuctioneerAction (src/Auctioneer.py:95-102) = HardcodedAuctioneerAcceptor
 * .selectAction per core (src/HardcodedModules.py:48-78) on the CURRENT state: the table index of
 * the best offeredReward/necessaryTime offer to each idle core (N*L = reject, also for cores the
 * auctioneer does not own).  random_ties != 0 picks uniformly among equal maxima (Philox, same
 * draw the step kernel would use this round), else the first.  out: int16 [B][C].  The step
 * kernel runs the same rule in-kernel when auctionMode != EXTERNAL; this entry point exists for
 * callers that want to see / override the auctioneer's action like the reference scripts do. */
int msched_auctioneer_action(void *handle, int random_ties, int16_t *out_dev, void *stream);

/* DividedHardcodedAgent.getActions of every agent (src/Agent.py:622-641) = SchedulingEnv.getActionForAllAgents
 * of HardcodedFixPriceEnvironment (src/SchedulingEnvironment.py:150-172, 439-456; BASELINE config 1): per
 * (agent, core) HardcodedAcceptor.selectAction (src/HardcodedModules.py:16-45: accept the best
 * offeredReward/necessaryTime if it beats the own job's priority/remainingLength, else reject) and per
 * (agent, slot) HardcodedOfferer.selectAction (src/HardcodedModules.py:81-109: offer to a core with the lowest
 * priority/remainingLength, never abstain), read from the dense observation record obs_dev (as written by
 * msched_observe_dense / msched_step_observe) and written into the acceptor idx and offer core fields of the
 * action record action_dev.  Ties: candidate floor(u * #candidates) in index order, u from Philox (counter =
 * global env, round, stream 3, unit) when random_ties != 0, the first candidate otherwise, or from
 * u_override_dev float32 [B][N*C + N*L] (parity tests).  ncand_dev (optional) int32 [B][N*C + N*L]: number of tie
 * candidates of each unit (0 where no choice was made). */
int msched_hardcoded_actions(void *handle, const int16_t *obs_dev, int random_ties, const float *u_override_dev,
                             int16_t *action_dev, int32_t *ncand_dev, void *stream);

/* The rollout loop of BASELINE config 1 (src/trainHC.py:60-95: getActionForAllAgents -> step, hard-coded agents and
 * auctioneer) on the device: n_steps x ( SchedulingEnv.step ; DividedHardcodedAgent.getActions on the NEW
 * observations ) in ONE launch of the multi-step fused kernel -- the agents' units of an environment are evaluated on
 * the observation tile in shared memory and their actions become the next step's action tile without leaving the SM.
 * action_dev: IN the action record of the first step (msched_hardcoded_actions on the current observations), OUT the
 * agents' actions for the step after the last (so launches chain).  result_dev [n_steps][padded_envs][result_words];
 * obs_dev as in msched_step_multi (every step with obs_every, else the last step's).  Tie draws as in
 * msched_hardcoded_actions (Philox at the round of the observations; random_ties == 0: first candidate).  Same
 * results, bit for bit, as the two-launch loop.  Fixed prices, in-kernel auction, device spawn draws. */
int msched_rollout_hardcoded(void *handle, int16_t *action_dev, int n_steps, uint32_t *result_dev, int16_t *obs_dev,
                             int obs_every, int random_ties, void *stream);

/* debug / parity: reference-shaped int32 dump of envs [env0, env0+count):
 * core [C][7] owner,prio,rem,jobid,kind,birth,init ; slot [N*L][7] prio,rem,jobid,kind,wait,
 * birth,init ; offer [N*L][5] core(0 none),recipient,price,time,offerID ;
 * chain [C][chainCap][5] offerer,recipient,price,time,round (newest first, -1 pad) ;
 * chain_len [C] ; misc [4] jobCounter, flags, 0, 0.  Device pointers, any may be NULL. */
int msched_export_state(void *handle, int env0, int count, int32_t *core, int32_t *slot,
                        int32_t *offer, int32_t *chain, int32_t *chain_len, int32_t *misc,
                        void *stream);

/* ---- policy side ----
 * ActorCritic.act (src/PPOmodules.py:53-63) for a group of identically shaped nets:
 * Linear(in,h)-Tanh-Linear(h,h)-Tanh-Linear(h,A)-Softmax, inverse-CDF categorical sample,
 * Categorical.log_prob.  weights: float32, torch layout, per net
 * [W1 h*in | b1 h | W2 h*h | b2 h | W3 A*h | b3 A]. */
typedef struct MschedMlpGroup {
    int32_t n_in, n_hidden, n_actions, n_nets;
    int32_t unit_div; /* unit u is served by net (u / unit_div) % n_nets: 1 = one net per unit
                         (divided) or one net for all (n_nets = 1, globally shared); C or L = one
                         net per agent (locally shared) */
    int32_t reserved;
    const float *weights; /* device, n_nets * msched_mlp_param_count() floats */
} MschedMlpGroup;

int msched_mlp_param_count(int n_in, int n_hidden, int n_actions);

/* Inputs/outputs of one actor launch over n_envs * units rows (row index r = b*units + u).
 * Unit u of env b reads its observation at x + b*env_stride + u*x_stride (int16 elements;
 * env_stride 0 means units*x_stride, i.e. a dense [M][x_stride] matrix) and is evaluated by net
 * (u / unit_div) % n_nets.  seed/step select the Philox stream (counter = (row_offset + r, step));
 * u_override float32 [M] replaces the draws (parity tests).  Any output may be NULL. */
typedef struct MschedActorIO {
    const int16_t *x;
    int32_t x_stride, units, n_envs, n_cores;
    int64_t env_stride, row_offset;
    uint64_t seed, step;
    const float *u_override;
    int32_t *action;         /* int32 [M] */
    float *logprob;          /* float32 [M] */
    float *probs;            /* float32 [M][A], tests only */
    int16_t *action_rec;     /* action of (b,u) also stored at action_rec[b*action_rec_stride + u]:
                                lets the kernel write straight into the env's action record */
    int64_t action_rec_stride;
    /* FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332): when gather_core != NULL the row
     * at x is the unit's OFFER observation [2C+2] and the net's 4 inputs are gathered from it by the
     * core chooser's action a = gather_core[r]: [row[2a], row[2a+1], row[2C], row[2C+1]], or
     * [-5,-5,-5,-5] for a == 0, in which case the reported action is -5 (quirk Q1). */
    const int32_t *gather_core;
    int16_t *x_used;         /* optional int16 [M][n_in]: the input actually fed (PPO buffer.states) */
    uint64_t *timeline;      /* diagnostics (NULL in production): the tensor-core kernel records 8 x uint64
                                clock64 stamps per CTA for its first tile */
    const uint64_t *step_dev; /* optional DEVICE step counter read instead of `step` (the caller advances
                                it, e.g. inside a CUDA graph that replays the rollout step) */
} MschedActorIO;

int msched_actor_forward(const MschedMlpGroup *nets, const MschedActorIO *io, void *stream);

/* ---- every PPO unit of a rollout step in ONE launch ----
 * The agents' getActions of one step (src/Agent.py:504-515 divided / shared fixed-price agents, :589-601 free-price
 * agents): PPO.selectAction (src/PPOmodules.py:114-125) of every acceptor unit and every offer unit -- for free
 * prices FreePriceOfferPPO.selectAction (src/PPOmodules.py:312-332), core chooser then price chooser -- straight on
 * the dense observation record.  16-wide nets (the divided, locally and globally shared agents).  Per group: the
 * nets, where unit u's row sits in the observation record (x_offset + u * x_stride, int16 elements, the row's first
 * logical value), where its action goes in the action record (rec_offset + u), and the experience-buffer slot of
 * this step (any of action / logprob / x_used may be NULL; x_used rows are x_used_stride int16 apart and hold the
 * observation row word-aligned: one leading pad value when x_offset is odd).  price.nets.weights == NULL: fixed
 * prices (the offer unit is OfferPPO alone).  Sampling draws: Philox4x32-10, key = the group's seed (the core
 * chooser's for both choosers), counter = (global env >> 1 lo, hi, step lo, 4 << 28 | step hi : 12 | unit : 16);
 * words 0,1 serve the even / odd environment's acceptor or core-chooser row, words 2,3 their price-chooser rows;
 * u = (x >> 8) * 2^-24.  env_offset (global index of env 0) must be even.  u_override float32 [n_envs][units]
 * replaces the draws (parity tests).  Unsupported net shapes return MSCHED_E_ARG: use msched_actor_forward.
 * Two kernels serve the call with the same contract: tcgen05 tensor cores (fp16 hi/lo operand pairs, fp32
 * accumulate; needs 0 < input_bound <= 511) and fp32 SIMT; MSCHED_POLICY_STEP_IMPL=tc|simt forces one. */
typedef struct MschedPolicyGroup {
    MschedMlpGroup nets;
    int32_t units, x_offset, x_stride, rec_offset;
    uint64_t seed;
    int32_t *action;        /* int32 [n_envs][units] */
    float *logprob;         /* float32 [n_envs][units] */
    int16_t *x_used;        /* int16 [n_envs][units][x_used_stride] */
    int32_t x_used_stride, reserved;
    const float *u_override;
    float *probs;           /* float32 [n_envs][units][n_actions], tests only */
} MschedPolicyGroup;

typedef struct MschedPolicyStep {
    const int16_t *obs;     /* dense observation record */
    int64_t obs_stride;     /* int16 per env */
    int32_t n_envs, n_cores;
    int16_t *action_rec;    /* the env's action record (or NULL) */
    int64_t action_rec_stride;
    int64_t env_offset;
    uint64_t step;
    const uint64_t *step_dev; /* optional device step counter (CUDA-graph replays) */
    MschedPolicyGroup acceptor, core, price;
    int32_t input_bound;      /* the caller's bound on |observation value| (priorities, prices, lengths); 1..511 lets the
                                 tensor-core kernel take the inputs as exact fp16 operands, 0 = unknown (fp32 SIMT kernel) */
    int32_t reserved;
} MschedPolicyStep;

int msched_policy_step(const MschedPolicyStep *ps, void *stream);

/* DQNEntity.selectAction (src/DQNmodules.py:34-76) for a group of Q-nets Linear(in,16)-Tanh-Linear(16,A)
 * (weights per net [W1 16*in | b1 16 | W2 A*16 | b2 A], torch layout): epsilon-greedy action per
 * (environment, unit) row -- a uniformly random action with probability epsilon (the caller evaluates the
 * schedule RUN_END + (RUN_START-RUN_END)*exp(-round/RUN_DECAY)), else the first arg-max of Q.  Uses the
 * x / stride / units / n_envs / seed / step / action / action_rec fields of io; u_override, if set, is
 * float32 [M][2] (exploration draw, random-action draw).  q_out: optional float32 [M][A]. */
int msched_dqn_param_count(int n_in, int n_actions);
int msched_dqn_select(const MschedMlpGroup *nets, const MschedActorIO *io, float epsilon, float *q_out,
                      void *stream);

/* optimize_model (src/DQNmodules.py:97-154) up to the optimizer step, for a group of Q-nets: for every net the gradient
 * of SmoothL1Loss(Q_policy(s)[a], gamma * max_a' Q_target(s') + r) (beta 1, mean over the batch) over `batch`
 * transitions, clamped to [-1, 1] like `param.grad.data.clamp_(-1, 1)`.  state / next_state: int16
 * [batch][n_nets][n_in]; action int32, reward float32: [batch][n_nets]; policy / target / grad: float32
 * [n_nets][msched_dqn_param_count]; loss (optional): float32 [n_nets].  Follow with msched_adam_step (the reference
 * uses torch.optim.Adam with its default learning rate, src/Agent.py:313-320).  Bit-reproducible (no atomics). */
typedef struct MschedDqnBatch {
    const float *policy, *target;
    int32_t n_in, n_hidden, n_actions, n_nets, batch, reserved;
    const int16_t *state, *next_state;
    const int32_t *action;
    const float *reward;
    float gamma, reserved2;
    float *grad, *loss;
} MschedDqnBatch;
int msched_dqn_grad(const MschedDqnBatch *b, void *stream);

/* PPO.update returns prologue (src/PPOmodules.py:128-137): G_t = r_t + gamma*G_{t+1} over
 * the whole buffer in float64, cast to float32, optional (G-mean)/(std_unbiased+1e-7) per
 * unit.  rewards/out: float32 [T][M] time-major. */
int msched_returns(const float *rewards, int T, int M, double gamma, int normalise, float *out,
                   void *stream);

/* ---- PPO.update on the device ----
 * One epoch's gradient of PPO.update (src/PPOmodules.py:139-174) for a group of identically shaped
 * ActorCritic nets (src/PPOmodules.py:25-51; 16 hidden neurons, n_in <= 64, n_actions <= 16), forward and
 * backward in one kernel: evaluate (log-prob of the stored action, entropy, V(s)), ratios, clipped surrogate,
 * value_coef * MseLoss (0.5) and -entropy_coef * entropy (0.01), mean over the net's samples, gradient with
 * respect to every actor and critic parameter.  Sample (tb, u), tb in [0, n_tb) (time x environment), unit
 * u in [0, units): observation int16 at x + tb*x_tb_stride + u*x_unit_stride, stored action / old log-prob /
 * normalised return at index tb*units + u.  Selected net s (n_sel of them) has id net_ids[s] and learns from
 * the units unit_ids[s*units_per_net .. +units_per_net) (every net the same number: divided nets have one
 * unit each, shared nets several, the CENTRALISATION_SAMPLE rule a subset, src/SchedulingEnvironment.py:
 * 314-329).  Rows net_ids[s] of grad_actor / grad_critic ([n_nets][msched_mlp_param_count], torch layout)
 * are overwritten; other rows are not touched.  stats (optional) receives per selected net
 * [mean -min(surr1,surr2), mean (V-G)^2, mean entropy, M].  Gradients are bit-reproducible (fixed-order
 * reduction, no atomics).  workspace: device scratch of msched_ppo_workspace_bytes(). */
typedef struct MschedPpoBatch {
    const float *actor_weights, *critic_weights; /* device [n_nets][param_count(n_in,16,n_actions / 1)] */
    int32_t n_in, n_hidden, n_actions, n_nets;
    const int16_t *x;
    int64_t x_tb_stride, x_unit_stride; /* int16 elements */
    const int32_t *action;              /* [n_tb][units] */
    const float *logprob_old, *returns; /* [n_tb][units] */
    int64_t n_tb;
    int32_t units, n_sel, units_per_net, reserved;
    const int32_t *net_ids;  /* device int32 [n_sel] */
    const int32_t *unit_ids; /* device int32 [n_sel][units_per_net] */
    float eps_clip, entropy_coef, value_coef, reserved2;
    float *grad_actor, *grad_critic;
    float *stats; /* device float32 [n_sel][4] or NULL */
    void *workspace;
    uint64_t workspace_bytes;
} MschedPpoBatch;

int msched_ppo_workspace_bytes(const MschedPpoBatch *b, uint64_t *bytes);
int msched_ppo_grad(const MschedPpoBatch *b, void *stream);

/* torch.optim.Adam single-tensor step (src/PPOmodules.py:100-105: one learning rate per head; no weight
 * decay, no amsgrad) over a flat float32 buffer, torch's operation order; step counts from 1. */
int msched_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                     double beta1, double beta2, double eps, int64_t step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSCHED_H */
