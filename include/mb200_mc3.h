/*
 * mb200_mc3.h -- C-ABI of the MC^3 shard coordinator (libmb200_mc3.so): one process per GPU,
 * the heated chains of an analysis dealt out over the processes, NCCL over NVLink for the only
 * two exchanges Metropolis coupling needs.
 *
 * What it replaces in the reference (file:line under /root/reference) -- the reference's only
 * distributed backend is MPI, and its hot-path-related traffic is exactly this:
 *
 *   mb200_mc3_create            chain -> process map of SetLocalChainsAndDataSplits
 *                               (src/mcmc.c:18331-18376: numLocalChains = numGlobalChains /
 *                               num_procs, contiguous blocks, src/mcmc.c:617-646) and the initial
 *                               chainId[] (SetChainIds, src/mcmc.c:17721)
 *   mb200_mc3_exchange_begin/_end   the per-swap-generation state exchange: the reference sends
 *                               myStateInfo[5] = {lnL, lnPrior, chainId, swapRan, 0} between the
 *                               two owners with MPI_Isend / MPI_Irecv / MPI_Waitall
 *                               (src/mcmc.c:831-852); here every process contributes
 *                               {lnL, lnPrior} of its chains to ONE ncclAllGather
 *                               (2 doubles x numLocalChains per rank), issued on its own stream
 *                               so that it overlaps the next generation's likelihood launches
 *   mb200_mc3_attempt_swaps     GetSwappers + AttemptSwap for every run (src/mcmc.c:5213-5246,
 *                               591-760, 16941-16958): lnR = (T_B - T_A)(lnL_A + lnPr_A) +
 *                               (T_A - T_B)(lnL_B + lnPr_B) (:718), heats swapped, states stay;
 *                               evaluated redundantly by every process from the gathered table
 *                               with the shared swap generator, so all processes agree without a
 *                               second message
 *   mb200_mc3_reduce_sum        the end-of-run marginal-likelihood reduce, MPI_Reduce (SUM, 1
 *                               double per run) (src/mcmc.c:17246, 17468) -> ncclReduce
 *
 * Random numbers: the reference picks the swap pair from `swapSeed`, identical on every process
 * (src/mcmc.c:5217-5218), and takes the acceptance draw from process A's chain generator
 * (src/mcmc.c:835).  Here BOTH come from the shared swap generator (the reference's Park-Miller
 * generator, RandomNumber src/utils.c:13802-13815), so the swap sequence is a function of the
 * seed and of the chains' lnL only -- independent of how many processes the chains are spread
 * over (the reference's MPI build reseeds every process, src/mcmc.c:2331, and cannot offer that).
 *
 * Every function returns 0 or a negative MB200_MC3_ERROR_* code.  world == 1 needs no NCCL.
 */
#ifndef MB200_MC3_H_
#define MB200_MC3_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_MC3_SUCCESS          0
#define MB200_MC3_ERROR_GENERAL   -1
#define MB200_MC3_ERROR_RANGE     -2
#define MB200_MC3_ERROR_NCCL      -3
#define MB200_MC3_ERROR_CUDA      -4
#define MB200_MC3_ERROR_PROTOCOL  -5   /* exchange_end without begin, swaps before the exchange ... */

#define MB200_MC3_ID_BYTES 128         /* sizeof(ncclUniqueId) */

typedef struct mb200_mc3 mb200_mc3;

typedef struct mb200_mc3_config
{
    int    rank, world;        /* this process, number of processes (one per GPU)                */
    int    device;             /* CUDA device of this process                                    */
    int    num_runs;           /* chainParams.numRuns                                            */
    int    chains_per_run;     /* chainParams.numChains                                          */
    int    num_swaps;          /* chainParams.numSwaps: swap attempts per run and swap generation */
    double chain_temp;         /* chainParams.chainTemp: T = 1 / (1 + chainTemp * id)  (:18963)   */
    long   swap_seed;          /* swapSeed (same on every process)                               */
    int    backend;            /* MB200_MC3_NCCL (default) or MB200_MC3_LOOPBACK                  */
} mb200_mc3_config;

#define MB200_MC3_NCCL      0
/* no transport: exchange_begin/_end only move the local rows (what a world of 1 does).  Lets a
 * caller that owns another transport (a test harness on gloo, MPI ...) fill the remote rows with
 * mb200_mc3_table() between begin and end; every other rule stays the coordinator's */
#define MB200_MC3_LOOPBACK  1

/* rank 0 creates the NCCL id and ships it to the others by whatever channel the launcher has
 * (torch.distributed broadcast, MPI_Bcast, a file); world == 1 passes NULL to create */
int mb200_mc3_unique_id (char id[MB200_MC3_ID_BYTES]);
int mb200_mc3_create    (const mb200_mc3_config *config, const char id[MB200_MC3_ID_BYTES], mb200_mc3 **out);
int mb200_mc3_destroy   (mb200_mc3 *mc);

/* ---- chain -> process map (global chain index g = run * chains_per_run + chain) ------------ */
int    mb200_mc3_local_chain_count (const mb200_mc3 *mc);
int    mb200_mc3_first_local_chain (const mb200_mc3 *mc);
int    mb200_mc3_owner             (const mb200_mc3 *mc, int global_chain);
int    mb200_mc3_chain_id          (const mb200_mc3 *mc, int global_chain);   /* heat id: id % chains_per_run == 0 is cold */
double mb200_mc3_temperature       (const mb200_mc3 *mc, int global_chain);   /* Temperature (chainId[g])                  */

/* ---- per swap generation ------------------------------------------------------------------- */
/* local_lnl / local_lnprior: curLnL[] / curLnPr[] of this process' chains, in local order.
 * begin() queues copy-in, all-gather and copy-out on the coordinator's own stream and returns at
 * once; end() waits for it.  With every swapper of the coming attempt local (or world == 1) no
 * collective is issued at all, as in the reference (src/mcmc.c:668). */
int mb200_mc3_exchange_begin (mb200_mc3 *mc, const double *local_lnl, const double *local_lnprior);
int mb200_mc3_exchange_end   (mb200_mc3 *mc);
/* the gathered table: [num_runs * chains_per_run][3] = {lnL, lnPrior, chainId}; writable (LOOPBACK backend) */
double *mb200_mc3_table (mb200_mc3 *mc);
/* num_swaps attempts per run; returns how many were accepted in *accepted (may be NULL).  Every
 * process calls it after exchange_end and reaches the same decisions. */
int mb200_mc3_attempt_swaps (mb200_mc3 *mc, int *accepted);
/* would the coming attempt_swaps touch chains of two different processes?  (pure function of the
 * swap generator's state; used to skip the collective) */
int mb200_mc3_next_swaps_cross_ranks (const mb200_mc3 *mc);

/* ---- end of run ----------------------------------------------------------------------------- */
int mb200_mc3_reduce_sum (mb200_mc3 *mc, double *values, int count, int root);   /* in place on root */
int mb200_mc3_barrier    (mb200_mc3 *mc);

/* ---- bookkeeping ---------------------------------------------------------------------------- */
/* swapInfo[run][i][j] of the reference (src/mcmc.c:751-753): upper triangle (i < j) accepted,
 * lower triangle attempted; out[num_runs * chains_per_run * chains_per_run] */
int       mb200_mc3_swap_info (const mb200_mc3 *mc, int *out);
long long mb200_mc3_collectives (const mb200_mc3 *mc);     /* all-gathers issued so far          */
/* FNV-1a over every (swapA, swapB, accepted) decision so far: equal on every process, and equal
 * for the same seed and the same chains whatever the number of processes */
unsigned long long mb200_mc3_decision_hash (const mb200_mc3 *mc);
/* per run: the same hash over that run's decisions, and how many of the run's swaps were decided by
 * another process without this one seeing them (co-resident pairs in a generation without a
 * collective); a process with missed == 0 holds the run's complete history */
unsigned long long mb200_mc3_run_hash   (const mb200_mc3 *mc, int run);
long long          mb200_mc3_run_missed (const mb200_mc3 *mc, int run);

#ifdef __cplusplus
}
#endif
#endif /* MB200_MC3_H_ */
