/*
 * msched_oracle.c -- CPU restatement of the marl-scheduling rollout hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs are the only callers.  The
 * product path (marl_scheduling_b200/csrc) never links, calls or falls back to it.
 *
 * Parity status: PINNED.  Every function below is checked bit-exactly (integer state,
 * rewards, observations, offer-ID tables) against traces produced by running the
 * unmodified reference in the build container (oracle/gen_golden.py -> tests/golden/),
 * including the six known-answer traces of SURVEY.md Appendix C.
 *
 * Style: deliberately literal.  One World object per environment, array-of-structs,
 * the offer list / ID tables / liability chains are materialised exactly as the
 * reference does, and every function cites the reference lines it follows
 * (paths relative to /root/reference/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MSOR_MAX_KINDS 16

enum { MODE_DIV_FIXED = 0, MODE_DIV_FREE_COMM = 1, MODE_DIV_FREE_NONCOMM = 2, MODE_AGG_FIXED = 3 };
enum { TIE_FIRST = 0, TIE_PHILOX = 1 };
enum { FLAG_CHAIN_OVERFLOW = 1, FLAG_COLLECTION_FULL = 2, FLAG_ACTION_RANGE = 4,
       FLAG_SPAWN_RANGE = 8 };

typedef struct {
    int32_t B, N, C, L, J;
    int32_t newJobs, rewardMultiplier, episodeLength;
    int32_t freePrices, rewardMode, chainCap, tieMode;
    int32_t prio[MSOR_MAX_KINDS], len[MSOR_MAX_KINDS], fix[MSOR_MAX_KINDS];
    double cumProb[MSOR_MAX_KINDS]; /* World.accProbabilities, src/world.py:220-222 */
    double netZeroOfferReward;
    uint64_t seed;
    int64_t envOffset; /* global index of env 0 (multi-GPU sharding) */
} MsorConfig;

/* src/world.py:79-104 */
typedef struct {
    int jobID, ownerID, priority, remainingLength, initialLength, empty, wait, birthDate, jobKind;
} Job;

/* src/world.py:27-37 */
typedef struct {
    int coreID, ownerID;
    Job job;
} Core;

/* src/world.py:156-196 */
typedef struct {
    int offerID, offererID, recipientID, coreID, queuePosition, jobID;
    int offeredReward, necessaryTime, round, prio1, jobKind;
} Offer;

typedef struct {
    int coreID, ownerID, jobID, generatedReward, round; /* jobTerminationInfo tuple */
    int prio, initLen, dwell;                           /* Verweilzeit record, src/world.py:350-357 */
    double dwellNorm;
} Termination;

typedef struct {
    Core *cores;           /* [C] */
    Job *collection;       /* [N][L] */
    int *freeSlots;        /* [N] numberOfFreeSlots */
    Offer *offers;         /* [NL] world.offers */
    int nOffers;
    Offer *chain;          /* [C][K] liabilityList, index 0 = newest (appendleft) */
    int *chainLen;         /* [C] */
    Termination *term;     /* [C] jobTerminationInfo */
    int nTerm;
    Offer *accepted;       /* [C] acceptedOffers */
    int nAccepted;
    int *ids;              /* [N][C][NL] env.correspondingOfferIDs */
    int *aucIds;           /* [C][NL] env.auctioneer_correspondingOfferIDs */
    int *obsAcc;           /* [N][C][3+2NL] */
    int *obsOff;           /* [N][L][2C+2] */
    int *obsAuc;           /* [C][3+2NL] */
    int *formerPrio, *formerLen; /* [C] src/SchedulingEnvironment.py:24-25,70-71 */
    int round;
    int jobIDCounter;      /* Job.IDCounter, per world here */
    int offerIDCounter;    /* Offer.offerID */
    uint32_t flags;
    int64_t terminationRevenues;
    /* episode statistics the train scripts derive from world.acceptedOffers / world.verweilzeiten
     * (src/trainPPO.py:172-227): per job kind [sum of accepted prices, #accepted,
     * sum of (dwell - 1), #terminated] */
    int32_t stats[4 * MSOR_MAX_KINDS];
} World;

typedef struct {
    MsorConfig cfg;
    World *w; /* [B] */
} Msor;

/* ------------------------------------------------------------------ Philox4x32-10
 * Published algorithm (Salmon et al., SC'11).  Used only where the build defines its own
 * randomness (device-side spawn / tie-break draws); the reference uses Python's MT19937,
 * which is replaced by recorded draws in parity mode. */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void msor_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    philox4x32_10(ctr, key, out);
}

/* stream tags of the build-defined RNG contract (DESIGN.md "Device randomness") */
enum { STREAM_SPAWN = 0, STREAM_TIE = 1 };

static void draw(const MsorConfig *cfg, int64_t env, uint32_t round, uint32_t stream, uint32_t a,
                 uint32_t b, uint32_t out[4])
{
    uint64_t g = (uint64_t)(cfg->envOffset + env);
    uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), round, (stream << 28) | (a << 12) | b };
    uint32_t key[2] = { (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32) };
    philox4x32_10(ctr, key, out);
}


/* ------------------------------------------------------------------ world objects */
static Job emptyJob(void) /* Job(0,0,0,True,None,None), src/world.py:83-93 */
{
    Job j;
    j.jobID = -1; j.ownerID = -1; j.priority = -1; j.remainingLength = -1;
    j.initialLength = -1; j.empty = 1; j.wait = 0; j.birthDate = -1; j.jobKind = -1;
    return j;
}

static Job newJob(World *w, int ownerID, int priority, int initialLength, int birthDate, int kind)
{ /* src/world.py:94-104 */
    Job j;
    j.jobID = w->jobIDCounter++;
    j.ownerID = ownerID; j.priority = priority;
    j.remainingLength = initialLength; j.initialLength = initialLength;
    j.empty = 0; j.wait = 0; j.birthDate = birthDate; j.jobKind = kind;
    return j;
}

/* JobCollection.insertJob, src/world.py:123-133 */
static void insertJob(const MsorConfig *cfg, World *w, int agentIdx, Job job)
{
    Job *col = w->collection + (size_t)agentIdx * cfg->L;
    if (w->freeSlots[agentIdx] > 0) {
        for (int s = 0; s < cfg->L; ++s) {
            if (col[s].empty) {
                col[s] = job;
                w->freeSlots[agentIdx] -= 1;
                break;
            }
        }
    } else {
        w->flags |= FLAG_COLLECTION_FULL; /* reference: raise "<str>" */
    }
}

/* JobCollection.removeAndReturnEntry, src/world.py:135-141 */
static Job removeAndReturnEntry(const MsorConfig *cfg, World *w, int agentIdx, int queuePosition)
{
    Job *col = w->collection + (size_t)agentIdx * cfg->L;
    Job result = col[queuePosition];
    col[queuePosition] = emptyJob();
    w->freeSlots[agentIdx] += 1;
    return result;
}

/* Core.dispatchNewJobAndReturnOldOne, src/world.py:61-76 */
static Job dispatchNewJobAndReturnOldOne(Core *core, Job newjob)
{
    Job old = core->job.empty ? emptyJob() : core->job;
    core->job = newjob;
    core->ownerID = newjob.empty ? 0 : newjob.ownerID;
    return old;
}

/* World.executeAnOffer, src/world.py:261-293 */
static void executeAnOffer(const MsorConfig *cfg, World *w, int offerID)
{
    Offer *offer = NULL;
    for (int k = 0; k < w->nOffers; ++k)
        if (w->offers[k].offerID == offerID) { offer = &w->offers[k]; break; }
    if (!offer) return; /* reference: StopIteration; unreachable through the tables */
    Core *core = &w->cores[offer->coreID - 1];
    if (offer->recipientID == core->ownerID) {
        int recipientID = offer->recipientID;
        Job nj = removeAndReturnEntry(cfg, w, offer->offererID - 1, offer->queuePosition);
        nj.wait = 0;
        Job old = dispatchNewJobAndReturnOldOne(core, nj);
        if (recipientID != 0)
            insertJob(cfg, w, recipientID - 1, old);
        /* liabilityList[core].appendleft(copy with round = now), :285-289 */
        int c = core->coreID - 1;
        Offer entry = *offer;
        entry.round = w->round;
        if (w->chainLen[c] < cfg->chainCap) {
            Offer *ch = w->chain + (size_t)c * cfg->chainCap;
            memmove(ch + 1, ch, sizeof(Offer) * (size_t)w->chainLen[c]);
            ch[0] = entry;
            w->chainLen[c] += 1;
        } else {
            w->flags |= FLAG_CHAIN_OVERFLOW; /* build-defined capacity, see DESIGN.md */
        }
        w->accepted[w->nAccepted++] = *offer; /* :293 */
        if (offer->jobKind >= 0) { /* prices.append((offer.offeredReward, offer.jobKind)), src/trainPPO.py:172-174 */
            w->stats[4 * offer->jobKind + 0] += offer->offeredReward;
            w->stats[4 * offer->jobKind + 1] += 1;
        }
    }
}

/* World.executeAgentAcceptions1, src/world.py:391-404 */
static void executeAgentAcceptions1(const MsorConfig *cfg, World *w, const int32_t *acc)
{
    const int NL = cfg->N * cfg->L;
    for (int i = 0; i < cfg->N; ++i)
        for (int j = 0; j < cfg->C; ++j) {
            int chosen = acc[i * cfg->C + j];
            if (chosen < NL) {
                if (chosen < 0) { w->flags |= FLAG_ACTION_RANGE; continue; } /* IndexError-ish */
                int id = w->ids[((size_t)i * cfg->C + j) * NL + chosen];
                if (0 < id) executeAnOffer(cfg, w, id);
            } else if (chosen != NL) {
                w->flags |= FLAG_ACTION_RANGE; /* reference: AssertionError */
            }
        }
}

/* World.executeAuctioneerAcceptions, src/world.py:378-389 */
static void executeAuctioneerAcceptions(const MsorConfig *cfg, World *w, const int32_t *auc)
{
    const int NL = cfg->N * cfg->L;
    for (int j = 0; j < cfg->C; ++j) {
        int chosen = auc[j];
        if (chosen < NL) {
            if (chosen < 0) { w->flags |= FLAG_ACTION_RANGE; continue; }
            int id = w->aucIds[(size_t)j * NL + chosen];
            if (0 < id) executeAnOffer(cfg, w, id);
        } else if (chosen != NL) {
            w->flags |= FLAG_ACTION_RANGE;
        }
    }
}

/* World.processOneTimestepAndUpdateOwnership, src/world.py:336-367 */
static void processOneTimestepAndUpdateOwnership(const MsorConfig *cfg, World *w)
{
    for (int c = 0; c < cfg->C; ++c) {
        Core *core = &w->cores[c];
        if (!core->job.empty) {
            core->job.remainingLength -= 1;
            if (core->job.remainingLength == 0) {
                Termination *t = &w->term[w->nTerm++];
                t->coreID = core->coreID;
                t->ownerID = core->ownerID;
                t->jobID = core->job.jobID;
                t->generatedReward = cfg->rewardMultiplier * core->job.priority;
                t->round = w->round + 1;
                t->prio = core->job.priority;
                t->initLen = core->job.initialLength;
                t->dwell = w->round - core->job.birthDate;
                t->dwellNorm = (double)(w->round - core->job.birthDate - 1) /
                               (double)core->job.initialLength;
                if (core->job.jobKind >= 0) { /* Verweilzeit record by (priority, length) = job kind */
                    w->stats[4 * core->job.jobKind + 2] += w->round - core->job.birthDate - 1;
                    w->stats[4 * core->job.jobKind + 3] += 1;
                }
                core->job = emptyJob(); /* assignCoreToAuctioneer, :57-59 */
                core->ownerID = 0;
            }
        }
    }
}

static void createOffers(const MsorConfig *cfg, World *w, const int32_t *offc, const int32_t *offp)
{ /* createFixPriceOfferObjectsFromActions :406-443 / createFreePrice... :445-478 */
    for (int i = 0; i < cfg->N; ++i)
        for (int j = 0; j < cfg->L; ++j) {
            int action = offc[i * cfg->L + j];
            int coreID = action + 1;
            Core *core = (coreID >= 1 && coreID <= cfg->C) ? &w->cores[coreID - 1] : NULL;
            Job *job = &w->collection[(size_t)i * cfg->L + j];
            int offeredReward;
            if (cfg->freePrices) {
                offeredReward = offp[i * cfg->L + j];
            } else { /* listOfFixPrices[jobKind]; jobKind -1 reads the last entry harmlessly */
                int k = job->jobKind < 0 ? cfg->J - 1 : job->jobKind;
                offeredReward = cfg->fix[k];
            }
            if (core != NULL && !job->empty && !job->wait) {
                Offer *o = &w->offers[w->nOffers++];
                o->offerID = w->offerIDCounter++;
                o->offererID = i + 1;
                o->recipientID = core->ownerID;
                o->coreID = core->coreID;
                o->queuePosition = j;
                o->jobID = job->jobID;
                o->offeredReward = offeredReward;
                o->necessaryTime = job->remainingLength;
                o->round = w->round;
                o->prio1 = job->priority;
                o->jobKind = job->jobKind;
                job->wait = 1;
            } else {
                job->wait = 0;
            }
        }
}

/* World.fillQueuesWithNewRandomJobs :369-376 + Agent.fillCollectionRandomly src/Agent.py:50-70 */
static void fillQueuesWithNewRandomJobs(const MsorConfig *cfg, World *w, int64_t env,
                                        const double *spawnU, const uint8_t *spawnKind)
{
    for (int i = 0; i < cfg->N; ++i) {
        int owned = 0;
        for (int c = 0; c < cfg->C; ++c) owned += (w->cores[c].ownerID == i + 1);
        if (owned + cfg->newJobs <= w->freeSlots[i]) {
            for (int k = 0; k < cfg->newJobs; ++k) {
                int kind = -1;
                if (spawnKind) {
                    kind = spawnKind[i * cfg->newJobs + k];
                } else {
                    double u;
                    if (spawnU) {
                        u = spawnU[i * cfg->newJobs + k];
                    } else {
                        /* draw d = agent*newJobs + k uses word d%4 of Philox call d/4 */
                        const int d = i * cfg->newJobs + k;
                        uint32_t x[4];
                        draw(cfg, env, (uint32_t)w->round, STREAM_SPAWN, (uint32_t)(d >> 2), 0u, x);
                        u = (double)x[d & 3] * (1.0 / 4294967296.0);
                    }
                    for (int q = 0; q < cfg->J; ++q)
                        if (u < cfg->cumProb[q]) { kind = q; break; }
                }
                if (kind < 0 || kind >= cfg->J) { /* reference: UnboundLocalError (Q13) */
                    w->flags |= FLAG_SPAWN_RANGE;
                    kind = cfg->J - 1;
                }
                insertJob(cfg, w, i, newJob(w, i + 1, cfg->prio[kind], cfg->len[kind], w->round, kind));
            }
        }
    }
}

/* DividedAgent.getAcceptorObservationTensorAndIDs src/Agent.py:167-212 (ownerID = agent)
 * getAuctioneerAcceptorObservationTensorAndIDs src/Auctioneer.py:34-77 (ownerID = 0) */
static void acceptorObservation(const MsorConfig *cfg, const World *w, int who, int coreID,
                                int *obs, int *ids)
{
    const int NL = cfg->N * cfg->L;
    const Core *core = &w->cores[coreID - 1];
    int own = (core->ownerID == who);
    obs[0] = own;
    obs[1] = own ? core->job.priority : -1;
    obs[2] = own ? core->job.remainingLength : -1;
    int n = 0;
    for (int k = 0; k < w->nOffers; ++k) {
        const Offer *o = &w->offers[k];
        if (o->recipientID == who && o->coreID == coreID) {
            if (n < NL) { /* Agent.pad truncates, Auctioneer.pad does not; n<=NL always */
                obs[3 + 2 * n] = o->offeredReward;
                obs[4 + 2 * n] = o->necessaryTime;
                ids[n] = o->offerID;
            }
            ++n;
        }
    }
    for (; n < NL; ++n) { obs[3 + 2 * n] = -2; obs[4 + 2 * n] = -2; ids[n] = -2; }
}

/* DividedAgent.getOfferNetObservationTensor src/Agent.py:271-300 */
static void offerObservation(const MsorConfig *cfg, const World *w, int agentIdx, int slot, int *obs)
{
    for (int c = 0; c < cfg->C; ++c) {
        obs[2 * c] = w->cores[c].job.priority;
        obs[2 * c + 1] = w->cores[c].job.remainingLength;
    }
    const Job *j = &w->collection[(size_t)agentIdx * cfg->L + slot];
    obs[2 * cfg->C] = j->priority;
    obs[2 * cfg->C + 1] = j->remainingLength;
}

/* the observation half of SchedulingEnv.step / reset, src/SchedulingEnvironment.py:44-58,85-109 */
static void gatherAllObservations(const MsorConfig *cfg, World *w)
{
    const int NL = cfg->N * cfg->L, W = 3 + 2 * NL;
    for (int i = 0; i < cfg->N; ++i) {
        for (int c = 0; c < cfg->C; ++c)
            acceptorObservation(cfg, w, i + 1, c + 1, w->obsAcc + ((size_t)i * cfg->C + c) * W,
                                w->ids + ((size_t)i * cfg->C + c) * NL);
        for (int s = 0; s < cfg->L; ++s)
            offerObservation(cfg, w, i, s, w->obsOff + ((size_t)i * cfg->L + s) * (2 * cfg->C + 2));
    }
    for (int c = 0; c < cfg->C; ++c)
        acceptorObservation(cfg, w, 0, c + 1, w->obsAuc + (size_t)c * W, w->aucIds + (size_t)c * NL);
}

/* HardcodedModules.calculateRewardRatio src/HardcodedModules.py:5-13 */
static double calculateRewardRatio(int priority, int remainingLength)
{
    if (priority == -1 || remainingLength == -1) return -1;
    if (priority == -2 || remainingLength == -2) return -1;
    return (double)priority / (double)remainingLength;
}

/* HardcodedAuctioneerAcceptor.selectAction src/HardcodedModules.py:54-78; the uniformly random
 * arg-max (random.sample) is replaced by tieMode: first candidate, or a Philox draw. */
static int auctioneerSelectAction(const MsorConfig *cfg, const World *w, int64_t env, int coreIdx)
{
    const int NL = cfg->N * cfg->L;
    const int *obs = w->obsAuc + (size_t)coreIdx * (3 + 2 * NL);
    if (obs[0] == 0) return NL;
    double own = calculateRewardRatio(obs[1], obs[2]);
    double best = -1e300;
    for (int k = 0; k < NL; ++k) {
        double r = calculateRewardRatio(obs[3 + 2 * k], obs[4 + 2 * k]);
        if (r > best) best = r;
    }
    if (!(best > own)) return NL;
    int ncand = 0;
    for (int k = 0; k < NL; ++k)
        if (calculateRewardRatio(obs[3 + 2 * k], obs[4 + 2 * k]) == best) ++ncand;
    int pick = 0;
    if (cfg->tieMode == TIE_PHILOX && ncand > 1) {
        uint32_t x[4];
        /* core j uses word j%4 of Philox call j/4 of the tie stream */
        draw(cfg, env, (uint32_t)w->round, STREAM_TIE, (uint32_t)(coreIdx >> 2), 0, x);
        pick = (int)(((uint64_t)x[coreIdx & 3] * (uint64_t)ncand) >> 32);
    }
    for (int k = 0; k < NL; ++k)
        if (calculateRewardRatio(obs[3 + 2 * k], obs[4 + 2 * k]) == best) {
            if (pick == 0) return k;
            --pick;
        }
    return NL;
}

/* SchedulingEnv.calculateAverageAcceptionQuality src/SchedulingEnvironment.py:174-192 */
static void acceptionQuality(const World *w, double *sum, int *cnt)
{
    double s = 0.0;
    int n = 0;
    for (int k = 0; k < w->nAccepted; ++k) {
        const Offer *o = &w->accepted[k];
        if (o->recipientID == 0) continue;
        double q = (double)o->offeredReward / (double)o->necessaryTime;
        int fp = w->formerPrio[o->coreID - 1];
        q -= (fp != -1) ? ((double)fp / (double)w->formerLen[o->coreID - 1]) : 0.0;
        q *= 10.0;
        s += q;
        ++n;
    }
    *sum = s;
    *cnt = n;
}

static int pyRound(double x) { return (int)nearbyint(x); } /* round(): half-to-even */

typedef struct {
    double *rOffer, *rPrice; /* [N][RL] */
    int64_t *rAcceptor;      /* [N][RC] */
    int64_t *rAuctioneer;    /* [C] */
    int64_t *rAgent;         /* [N] */
} RewardOut;

/* Reward.py: getDividedFixedPricesReward :146-212, getDividedFreePricesReward :6-89,
 * getAggregatedFixedPricesReward :92-143 */
static void getRewards(const MsorConfig *cfg, World *w, RewardOut r)
{
    const int N = cfg->N, C = cfg->C, L = cfg->L;
    const int agg = cfg->rewardMode == MODE_AGG_FIXED;
    const int freeM = cfg->rewardMode == MODE_DIV_FREE_COMM || cfg->rewardMode == MODE_DIV_FREE_NONCOMM;
    const int RL = agg ? 1 : L, RC = agg ? 1 : C;
    for (int k = 0; k < N * RL; ++k) { r.rOffer[k] = 0; r.rPrice[k] = 0; }
    for (int k = 0; k < N * RC; ++k) r.rAcceptor[k] = 0;
    for (int k = 0; k < C; ++k) r.rAuctioneer[k] = 0;
    for (int k = 0; k < N; ++k) r.rAgent[k] = 0;

    for (int k = 0; k < w->nAccepted; ++k) {
        const Offer *o = &w->accepted[k];
        int a = o->offererID - 1, slot = o->queuePosition;
        if (agg) {
            r.rOffer[a] += o->prio1; /* :103 */
        } else if (!freeM) {
            r.rOffer[a * L + slot] = o->prio1; /* :170 */
        } else {
            double price;
            int d = o->prio1 - o->offeredReward;
            if (cfg->rewardMode == MODE_DIV_FREE_COMM)
                price = (d == 0) ? cfg->netZeroOfferReward : (double)d; /* :29-33 */
            else
                price = (d >= 0) ? (double)o->prio1 : (double)d; /* :43-47 */
            r.rPrice[a * L + slot] = price;
            r.rOffer[a * L + slot] = o->prio1;
        }
    }
    for (int t = 0; t < w->nTerm; ++t) {
        const Termination *T = &w->term[t];
        int c = T->coreID - 1, owner = T->ownerID - 1, R = T->generatedReward;
        if (agg) {
            r.rAcceptor[owner] += R; /* :122 */
            r.rAgent[owner] += R;
        } else {
            r.rAcceptor[owner * C + c] = R; /* :63 / :191 */
            if (!freeM) {
                r.rAgent[owner] += R; /* :192 (not in the free-price variant, Q6) */
                w->terminationRevenues += R;
            }
        }
        int last = T->round, timeMeasure = 0;
        const Offer *ch = w->chain + (size_t)c * cfg->chainCap;
        for (int e = 0; e < w->chainLen[c]; ++e) {
            const Offer *en = &ch[e];
            timeMeasure += last - en->round;
            last = en->round;
            double ratio = (double)en->offeredReward / (double)en->necessaryTime;
            int traded = pyRound(ratio * (double)timeMeasure);
            if (agg) {
                r.rAcceptor[en->offererID - 1] -= traded;
                r.rAgent[en->offererID - 1] -= traded;
                if (en->recipientID > 0) r.rAgent[en->recipientID - 1] += traded;
            } else {
                r.rAcceptor[(en->offererID - 1) * C + c] -= traded;
                r.rAgent[en->offererID - 1] -= traded;
                if (en->recipientID > 0) {
                    r.rAcceptor[(en->recipientID - 1) * C + c] += traded;
                    r.rAgent[en->recipientID - 1] += traded;
                }
            }
            if (en->recipientID == 0) r.rAuctioneer[c] = traded;
        }
        w->chainLen[c] = 0; /* resetLiabilityListForACore */
    }
}

/* ------------------------------------------------------------------ public API */
static size_t worldInts(const MsorConfig *c) { (void)c; return 0; }

void *msor_create(const MsorConfig *cfg)
{
    (void)worldInts;
    if (cfg->J < 1 || cfg->J > MSOR_MAX_KINDS || cfg->B < 0) return NULL;
    Msor *h = (Msor *)calloc(1, sizeof(Msor));
    h->cfg = *cfg;
    const int N = cfg->N, C = cfg->C, L = cfg->L, NL = N * L, W = 3 + 2 * NL;
    h->w = (World *)calloc((size_t)(cfg->B > 0 ? cfg->B : 1), sizeof(World));
    for (int b = 0; b < cfg->B; ++b) {
        World *w = &h->w[b];
        w->cores = (Core *)calloc((size_t)C, sizeof(Core));
        w->collection = (Job *)calloc((size_t)NL, sizeof(Job));
        w->freeSlots = (int *)calloc((size_t)N, sizeof(int));
        w->offers = (Offer *)calloc((size_t)NL, sizeof(Offer));
        w->chain = (Offer *)calloc((size_t)C * cfg->chainCap, sizeof(Offer));
        w->chainLen = (int *)calloc((size_t)C, sizeof(int));
        w->term = (Termination *)calloc((size_t)C, sizeof(Termination));
        w->accepted = (Offer *)calloc((size_t)C, sizeof(Offer));
        w->ids = (int *)calloc((size_t)N * C * NL, sizeof(int));
        w->aucIds = (int *)calloc((size_t)C * NL, sizeof(int));
        w->obsAcc = (int *)calloc((size_t)N * C * W, sizeof(int));
        w->obsOff = (int *)calloc((size_t)NL * (2 * C + 2), sizeof(int));
        w->obsAuc = (int *)calloc((size_t)C * W, sizeof(int));
        w->formerPrio = (int *)calloc((size_t)C, sizeof(int));
        w->formerLen = (int *)calloc((size_t)C, sizeof(int));
    }
    return h;
}

void msor_destroy(void *hv)
{
    Msor *h = (Msor *)hv;
    if (!h) return;
    for (int b = 0; b < h->cfg.B; ++b) {
        World *w = &h->w[b];
        free(w->cores); free(w->collection); free(w->freeSlots); free(w->offers); free(w->chain);
        free(w->chainLen); free(w->term); free(w->accepted); free(w->ids); free(w->aucIds);
        free(w->obsAcc); free(w->obsOff); free(w->obsAuc); free(w->formerPrio); free(w->formerLen);
    }
    free(h->w);
    free(h);
}

/* World.__init__ src/world.py:210-254 + SchedulingEnv.__init__/reset src/SchedulingEnvironment.py:22-30,85-109 */
void msor_reset(void *hv)
{
    Msor *h = (Msor *)hv;
    const MsorConfig *cfg = &h->cfg;
    for (int b = 0; b < cfg->B; ++b) {
        World *w = &h->w[b];
        for (int c = 0; c < cfg->C; ++c) {
            w->cores[c].coreID = c + 1;
            w->cores[c].ownerID = 0;
            w->cores[c].job = emptyJob();
            w->chainLen[c] = 0;
            w->formerPrio[c] = -1;
            w->formerLen[c] = -1;
        }
        for (int k = 0; k < cfg->N * cfg->L; ++k) w->collection[k] = emptyJob();
        for (int i = 0; i < cfg->N; ++i) w->freeSlots[i] = cfg->L;
        w->nOffers = 0; w->nTerm = 0; w->nAccepted = 0;
        w->round = 0; w->jobIDCounter = 1; w->offerIDCounter = 1;
        w->flags = 0; w->terminationRevenues = 0;
        memset(w->stats, 0, sizeof(w->stats));
        gatherAllObservations(cfg, w);
    }
}

/* One SchedulingEnv.step for envs [b0,b1).  src/SchedulingEnvironment.py:32-83 and
 * World.step1 src/world.py:295-334.  All arrays are env-major.  auc == NULL runs the
 * hard-coded auctioneer (Auctioneer.getAuctioneerAction src/Auctioneer.py:95-102). */
void msor_step_range(void *hv, int b0, int b1, const int32_t *offc, const int32_t *offp,
                     const int32_t *acc, const int32_t *auc, const double *spawnU,
                     const uint8_t *spawnKind, double *rOffer, double *rPrice,
                     int64_t *rAcceptor, int64_t *rAuctioneer, int64_t *rAgent,
                     double *qualitySum, int32_t *qualityCnt, uint8_t *done, int32_t *aucOut,
                     int32_t *nAccepted, int32_t *nTerm, uint32_t *flags)
{
    Msor *h = (Msor *)hv;
    const MsorConfig *cfg = &h->cfg;
    const int N = cfg->N, C = cfg->C, L = cfg->L;
    const int agg = cfg->rewardMode == MODE_AGG_FIXED;
    const int RL = agg ? 1 : L, RC = agg ? 1 : C;
    int32_t *aucTmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)C);
    for (int b = b0; b < b1; ++b) {
        World *w = &h->w[b];
        const int32_t *a = auc ? auc + (size_t)b * C : NULL;
        if (!a) {
            for (int c = 0; c < C; ++c) aucTmp[c] = auctioneerSelectAction(cfg, w, b, c);
            a = aucTmp;
        }
        if (aucOut) memcpy(aucOut + (size_t)b * C, a, sizeof(int32_t) * (size_t)C);
        /* ---- World.step1 ---- */
        for (int t = 0; t < w->nTerm; ++t) w->chainLen[w->term[t].coreID - 1] = 0; /* :309-310 */
        w->nTerm = 0;
        w->nAccepted = 0;
        executeAgentAcceptions1(cfg, w, acc + (size_t)b * N * C);
        executeAuctioneerAcceptions(cfg, w, a);
        processOneTimestepAndUpdateOwnership(cfg, w);
        w->nOffers = 0;
        w->offerIDCounter = 1;
        createOffers(cfg, w, offc + (size_t)b * N * L, offp ? offp + (size_t)b * N * L : NULL);
        fillQueuesWithNewRandomJobs(cfg, w, b,
                                    spawnU ? spawnU + (size_t)b * N * cfg->newJobs : NULL,
                                    spawnKind ? spawnKind + (size_t)b * N * cfg->newJobs : NULL);
        w->round += 1;
        /* ---- SchedulingEnv.step tail ---- */
        gatherAllObservations(cfg, w);
        double qs; int qc;
        acceptionQuality(w, &qs, &qc);
        if (qualitySum) qualitySum[b] = qs;
        if (qualityCnt) qualityCnt[b] = qc;
        RewardOut r;
        r.rOffer = rOffer + (size_t)b * N * RL;
        r.rPrice = rPrice + (size_t)b * N * RL;
        r.rAcceptor = rAcceptor + (size_t)b * N * RC;
        r.rAuctioneer = rAuctioneer + (size_t)b * C;
        r.rAgent = rAgent + (size_t)b * N;
        getRewards(cfg, w, r);
        if (done) done[b] = (uint8_t)((w->round % cfg->episodeLength) == 0);
        for (int c = 0; c < C; ++c) {
            w->formerPrio[c] = w->cores[c].job.priority;
            w->formerLen[c] = w->cores[c].job.remainingLength;
        }
        if (nAccepted) nAccepted[b] = w->nAccepted;
        if (nTerm) nTerm[b] = w->nTerm;
        if (flags) flags[b] = w->flags;
    }
    free(aucTmp);
}

/* dense observations of env b as int32: acc [N][C][3+2NL], off [N][L][2C+2], auc [C][3+2NL],
 * ids [N][C][NL], aucIds [C][NL]; any pointer may be NULL */
void msor_observe(void *hv, int b, int32_t *obsAcc, int32_t *obsOff, int32_t *obsAuc, int32_t *ids,
                  int32_t *aucIds)
{
    Msor *h = (Msor *)hv;
    const MsorConfig *cfg = &h->cfg;
    const World *w = &h->w[b];
    const int N = cfg->N, C = cfg->C, L = cfg->L, NL = N * L, W = 3 + 2 * NL;
    if (obsAcc) memcpy(obsAcc, w->obsAcc, sizeof(int) * (size_t)N * C * W);
    if (obsOff) memcpy(obsOff, w->obsOff, sizeof(int) * (size_t)NL * (2 * C + 2));
    if (obsAuc) memcpy(obsAuc, w->obsAuc, sizeof(int) * (size_t)C * W);
    if (ids) memcpy(ids, w->ids, sizeof(int) * (size_t)N * C * NL);
    if (aucIds) memcpy(aucIds, w->aucIds, sizeof(int) * (size_t)C * NL);
}

/* reference-shaped dump of env b.
 * core [C][7]  = owner, prio, rem, jobid, kind, birth, init
 * slot [NL][7] = prio, rem, jobid, kind, wait, birth, init
 * off  [NL][5] = core(0 = none), recip, price, time, id
 * chain [C][K][5] = offerer, recipient, price, time, round (newest first, -1 padded); chainLen [C]
 * accepted [C][5] = offerer, recipient, core, slot, price ; term [C][8] = core, owner, jobid, R,
 * round, prio, initLen, dwell ; termNorm [C] ; misc [6] = round, jobIDCounter, nAccepted, nTerm,
 * flags, terminationRevenues(lo32) */
void msor_export(void *hv, int b, int32_t *core, int32_t *slot, int32_t *off, int32_t *chain,
                 int32_t *chainLen, int32_t *accepted, int32_t *term, double *termNorm,
                 int32_t *misc)
{
    Msor *h = (Msor *)hv;
    const MsorConfig *cfg = &h->cfg;
    const World *w = &h->w[b];
    const int N = cfg->N, C = cfg->C, L = cfg->L, NL = N * L, K = cfg->chainCap;
    for (int c = 0; c < C; ++c) {
        const Job *j = &w->cores[c].job;
        int32_t *o = core + c * 7;
        o[0] = w->cores[c].ownerID; o[1] = j->priority; o[2] = j->remainingLength; o[3] = j->jobID;
        o[4] = j->jobKind; o[5] = j->birthDate; o[6] = j->initialLength;
    }
    for (int s = 0; s < NL; ++s) {
        const Job *j = &w->collection[s];
        int32_t *o = slot + s * 7;
        o[0] = j->priority; o[1] = j->remainingLength; o[2] = j->jobID; o[3] = j->jobKind;
        o[4] = j->wait; o[5] = j->birthDate; o[6] = j->initialLength;
        int32_t *f = off + s * 5;
        f[0] = 0; f[1] = f[2] = f[3] = f[4] = -1;
    }
    for (int k = 0; k < w->nOffers; ++k) {
        const Offer *o = &w->offers[k];
        int32_t *f = off + ((o->offererID - 1) * L + o->queuePosition) * 5;
        f[0] = o->coreID; f[1] = o->recipientID; f[2] = o->offeredReward; f[3] = o->necessaryTime;
        f[4] = o->offerID;
    }
    for (int c = 0; c < C; ++c) {
        chainLen[c] = w->chainLen[c];
        for (int e = 0; e < K; ++e) {
            int32_t *o = chain + ((size_t)c * K + e) * 5;
            if (e < w->chainLen[c]) {
                const Offer *en = &w->chain[(size_t)c * K + e];
                o[0] = en->offererID; o[1] = en->recipientID; o[2] = en->offeredReward;
                o[3] = en->necessaryTime; o[4] = en->round;
            } else {
                o[0] = o[1] = o[2] = o[3] = o[4] = -1;
            }
        }
    }
    for (int k = 0; k < C; ++k) {
        int32_t *a = accepted + k * 5;
        int32_t *t = term + k * 8;
        if (k < w->nAccepted) {
            const Offer *o = &w->accepted[k];
            a[0] = o->offererID; a[1] = o->recipientID; a[2] = o->coreID; a[3] = o->queuePosition;
            a[4] = o->offeredReward;
        } else {
            a[0] = a[1] = a[2] = a[3] = a[4] = -1;
        }
        if (k < w->nTerm) {
            const Termination *T = &w->term[k];
            t[0] = T->coreID; t[1] = T->ownerID; t[2] = T->jobID; t[3] = T->generatedReward;
            t[4] = T->round; t[5] = T->prio; t[6] = T->initLen; t[7] = T->dwell;
            termNorm[k] = T->dwellNorm;
        } else {
            for (int q = 0; q < 8; ++q) t[q] = -1;
            termNorm[k] = 0.0;
        }
    }
    misc[0] = w->round; misc[1] = w->jobIDCounter; misc[2] = w->nAccepted; misc[3] = w->nTerm;
    misc[4] = (int32_t)w->flags; misc[5] = (int32_t)w->terminationRevenues;
}

/* ------------------------------------------------------------------ PPO pieces
 * PPO.update returns prologue, src/PPOmodules.py:128-137: Python float64 reverse accumulation,
 * cast to float32, then (G - mean) / (std_unbiased + 1e-7) in float32.
 * rewards/out are [T][M] (time-major, M independent units). */
void msor_returns(const double *rewards, int T, int M, double gamma, int normalise, float *out)
{
    for (int m = 0; m < M; ++m) {
        double disc = 0.0;
        for (int t = T - 1; t >= 0; --t) {
            disc = rewards[(size_t)t * M + m] + gamma * disc;
            out[(size_t)t * M + m] = (float)disc;
        }
        if (normalise) {
            /* torch float32 mean / std; accumulate in double and round once (torch's own
             * summation order differs in the last ulp; tests allow 1e-5 relative) */
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += out[(size_t)t * M + m];
            float mean = (float)(s / T);
            double v = 0.0;
            for (int t = 0; t < T; ++t) {
                double d = (double)out[(size_t)t * M + m] - (double)mean;
                v += d * d;
            }
            float sd = (float)sqrt(v / (T - 1));
            for (int t = 0; t < T; ++t)
                out[(size_t)t * M + m] = (out[(size_t)t * M + m] - mean) / (sd + 1e-7f);
        }
    }
}

/* ActorCritic.actor / .critic forward, src/PPOmodules.py:32-48,53-72: Linear-Tanh-Linear-Tanh-
 * Linear(-Softmax) in float32.  x [M][nin]; weights are torch layout W[out][in].
 * probs [M][A] (actor, softmax=1) or values [M][1] (critic, A=1, softmax=0).
 * If u != NULL also samples action = first a with cdf(a) > u*sum (inverse CDF on the float32
 * probabilities) and returns log_prob as torch.distributions.Categorical computes it:
 * probs renormalised by their sum, clamped to [eps, 1-eps], then log. */
void msor_mlp_forward(const float *x, int M, int nin, int h, int A, const float *W1, const float *b1,
                      const float *W2, const float *b2, const float *W3, const float *b3,
                      int softmax, float *out, const float *u, int32_t *action, float *logprob)
{
    float *h1 = (float *)malloc(sizeof(float) * (size_t)h);
    float *h2 = (float *)malloc(sizeof(float) * (size_t)h);
    for (int m = 0; m < M; ++m) {
        const float *xm = x + (size_t)m * nin;
        for (int o = 0; o < h; ++o) {
            float s = b1[o];
            for (int k = 0; k < nin; ++k) s += W1[(size_t)o * nin + k] * xm[k];
            h1[o] = tanhf(s);
        }
        for (int o = 0; o < h; ++o) {
            float s = b2[o];
            for (int k = 0; k < h; ++k) s += W2[(size_t)o * h + k] * h1[k];
            h2[o] = tanhf(s);
        }
        float *om = out + (size_t)m * A;
        float mx = -INFINITY;
        for (int o = 0; o < A; ++o) {
            float s = b3[o];
            for (int k = 0; k < h; ++k) s += W3[(size_t)o * h + k] * h2[k];
            om[o] = s;
            if (s > mx) mx = s;
        }
        if (softmax) {
            float sum = 0.f;
            for (int o = 0; o < A; ++o) { om[o] = expf(om[o] - mx); sum += om[o]; }
            for (int o = 0; o < A; ++o) om[o] /= sum;
            if (u) {
                float tot = 0.f;
                for (int o = 0; o < A; ++o) tot += om[o];
                float thr = u[m] * tot, cdf = 0.f;
                int a = A - 1;
                for (int o = 0; o < A; ++o) {
                    cdf += om[o];
                    if (cdf > thr) { a = o; break; }
                }
                action[m] = a;
                float p = om[a] / tot;
                const float eps = 1.1920928955078125e-07f; /* torch.finfo(float32).eps */
                if (p < eps) p = eps;
                if (p > 1.f - eps) p = 1.f - eps;
                logprob[m] = logf(p);
            }
        }
    }
    free(h1);
    free(h2);
}

/* episode statistics of every env: out [B][J][4] int32 (see World.stats) */
void msor_stats(void *handle, int32_t *out)
{
    Msor *h = (Msor *)handle;
    const int J = h->cfg.J;
    for (int b = 0; b < h->cfg.B; ++b)
        memcpy(out + (size_t)b * 4 * J, h->w[b].stats, sizeof(int32_t) * 4 * (size_t)J);
}
