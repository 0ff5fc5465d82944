/*
 * mb200_mc3.c -- MC^3 shard coordinator (include/mb200_mc3.h): chain -> process map, the
 * per-swap-generation {lnL, lnPrior, chainId} exchange over NCCL, the swap decisions, the
 * end-of-run reduce.  Host C; the only device work is the collective itself (24 bytes per chain).
 *
 * The arithmetic follows the reference: GetSwappers (src/mcmc.c:5213-5246), AttemptSwap
 * (src/mcmc.c:591-760), Temperature (src/mcmc.c:18963-18975), RandomNumber
 * (src/utils.c:13802-13815).  Nothing is copied: the rules are restated on a gathered table
 * instead of pairwise messages.
 */
#include "mb200_mc3.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ROW 3                           /* doubles per chain in the table: lnL, lnPrior, chainId */

struct mb200_mc3
{
    mb200_mc3_config cfg;
    int         nGlobal, nLocal, first;         /* chains: all, mine, my first global index        */
    long        swapSeed;                       /* shared generator state                          */
    double     *table;                          /* [nGlobal][ROW] as of the last exchange          */
    int        *chainId;                        /* [nGlobal] heat ids as this process knows them   */
    int        *swapInfo;                       /* [runs][chains][chains]                          */
    unsigned long long *runHash;                /* [runs] FNV-1a over that run's decisions         */
    long long  *runMissed;                      /* [runs] swaps of the run decided elsewhere       */
    unsigned long long hash;
    long long   collectives;
    int         fresh;                          /* the table holds every chain's current row       */
    int         pending;                        /* exchange_begin issued, end not yet              */
    int         pendingCollective;
    /* transport */
    ncclComm_t   comm;
    cudaStream_t stream;
    cudaEvent_t  done;
    double      *hSend, *hRecv;                 /* pinned                                          */
    double      *dSend, *dRecv;
    double      *dRed;                          /* reduce scratch                                  */
    int          redCap;
};

static double NextRandom (long *seed)           /* RandomNumber, src/utils.c:13802-13815 */
{
    long lo, hi, test;

    hi = (*seed) / 127773;
    lo = (*seed) % 127773;
    test = 16807 * lo - 2836 * hi;
    if (test > 0)
        *seed = test;
    else
        *seed = test + 2147483647;
    return ((double)(*seed) / (double)2147483647);
}

static void PickSwappers (const mb200_mc3 *mc, long *seed, int run, int *a, int *b)   /* GetSwappers, random pairs */
{
    const int n = mc->cfg.chains_per_run;

    *a = (int) (NextRandom (seed) * n);
    *b = (int) (NextRandom (seed) * (n - 1));
    if (*b == *a)
        *b = n - 1;
    *a += run * n;
    *b += run * n;
}

static int OwnerOf (const mb200_mc3 *mc, int g)
{
    return g / mc->nLocal;
}

static double TemperatureOf (const mb200_mc3 *mc, int id)
{
    id %= mc->cfg.chains_per_run;
    return 1.0 / (1.0 + mc->cfg.chain_temp * id);
}

int mb200_mc3_unique_id (char id[MB200_MC3_ID_BYTES])
{
    ncclUniqueId u;
    if (sizeof(u) != MB200_MC3_ID_BYTES)
        return MB200_MC3_ERROR_GENERAL;
    if (ncclGetUniqueId (&u) != ncclSuccess)
        return MB200_MC3_ERROR_NCCL;
    memcpy (id, &u, MB200_MC3_ID_BYTES);
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_destroy (mb200_mc3 *mc)
{
    if (!mc)
        return MB200_MC3_SUCCESS;
    if (mc->stream)
        {
        cudaSetDevice (mc->cfg.device);
        cudaStreamSynchronize (mc->stream);
        }
    if (mc->comm)   ncclCommDestroy (mc->comm);
    if (mc->done)   cudaEventDestroy (mc->done);
    if (mc->hSend)  cudaFreeHost (mc->hSend);
    if (mc->hRecv)  cudaFreeHost (mc->hRecv);
    if (mc->dSend)  cudaFree (mc->dSend);
    if (mc->dRecv)  cudaFree (mc->dRecv);
    if (mc->dRed)   cudaFree (mc->dRed);
    if (mc->stream) cudaStreamDestroy (mc->stream);
    free (mc->table); free (mc->chainId); free (mc->swapInfo); free (mc->runHash); free (mc->runMissed);
    free (mc);
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_create (const mb200_mc3_config *cfg, const char id[MB200_MC3_ID_BYTES], mb200_mc3 **out)
{
    mb200_mc3 *mc;
    int        g, nGlobal;

    if (!cfg || !out)
        return MB200_MC3_ERROR_GENERAL;
    *out = NULL;
    nGlobal = cfg->num_runs * cfg->chains_per_run;
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world || cfg->num_runs < 1 || cfg->chains_per_run < 1 ||
        cfg->num_swaps < 0 || cfg->swap_seed <= 0)
        return MB200_MC3_ERROR_RANGE;
    /* SetLocalChainsAndDataSplits (src/mcmc.c:18336-18352): the chains must divide evenly, at least one per process */
    if (cfg->world > nGlobal || nGlobal % cfg->world != 0)
        return MB200_MC3_ERROR_RANGE;
    mc = (mb200_mc3 *) calloc (1, sizeof(mb200_mc3));
    if (!mc)
        return MB200_MC3_ERROR_GENERAL;
    mc->cfg = *cfg;
    mc->nGlobal = nGlobal;
    mc->nLocal  = nGlobal / cfg->world;
    mc->first   = cfg->rank * mc->nLocal;
    mc->swapSeed = cfg->swap_seed;
    mc->hash = 1469598103934665603ULL;
    mc->table     = (double *) calloc ((size_t)nGlobal * ROW, sizeof(double));
    mc->chainId   = (int *)    calloc ((size_t)nGlobal, sizeof(int));
    mc->swapInfo  = (int *)    calloc ((size_t)cfg->num_runs * cfg->chains_per_run * cfg->chains_per_run, sizeof(int));
    mc->runHash   = (unsigned long long *) calloc ((size_t)cfg->num_runs, sizeof(unsigned long long));
    mc->runMissed = (long long *) calloc ((size_t)cfg->num_runs, sizeof(long long));
    if (!mc->table || !mc->chainId || !mc->swapInfo || !mc->runHash || !mc->runMissed)
        { mb200_mc3_destroy (mc); return MB200_MC3_ERROR_GENERAL; }
    for (g=0; g<nGlobal; g++)
        {
        mc->chainId[g] = g;                      /* SetChainIds (src/mcmc.c:17721): the global chain number */
        mc->table[(size_t)g*ROW + 2] = (double) g;
        }
    for (g=0; g<cfg->num_runs; g++)
        mc->runHash[g] = 1469598103934665603ULL;

    if (cfg->world > 1 && cfg->backend == MB200_MC3_NCCL)
        {
        ncclUniqueId u;
        if (!id)
            { mb200_mc3_destroy (mc); return MB200_MC3_ERROR_RANGE; }
        memcpy (&u, id, MB200_MC3_ID_BYTES);
        if (cudaSetDevice (cfg->device) != cudaSuccess ||
            cudaStreamCreateWithFlags (&mc->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags (&mc->done, cudaEventDisableTiming) != cudaSuccess ||
            cudaMallocHost ((void **)&mc->hSend, (size_t)mc->nLocal * ROW * sizeof(double)) != cudaSuccess ||
            cudaMallocHost ((void **)&mc->hRecv, (size_t)nGlobal * ROW * sizeof(double)) != cudaSuccess ||
            cudaMalloc ((void **)&mc->dSend, (size_t)mc->nLocal * ROW * sizeof(double)) != cudaSuccess ||
            cudaMalloc ((void **)&mc->dRecv, (size_t)nGlobal * ROW * sizeof(double)) != cudaSuccess)
            { mb200_mc3_destroy (mc); return MB200_MC3_ERROR_CUDA; }
        if (ncclCommInitRank (&mc->comm, cfg->world, u, cfg->rank) != ncclSuccess)
            { mc->comm = NULL; mb200_mc3_destroy (mc); return MB200_MC3_ERROR_NCCL; }
        /* first use of a communicator sets up its channels (hundreds of microseconds): do it here,
           not inside somebody's timed swap generation */
        memset (mc->hSend, 0, (size_t)mc->nLocal * ROW * sizeof(double));
        for (g=0; g<2; g++)
            {
            if (cudaMemcpyAsync (mc->dSend, mc->hSend, (size_t)mc->nLocal * ROW * sizeof(double), cudaMemcpyHostToDevice, mc->stream) != cudaSuccess ||
                ncclAllGather (mc->dSend, mc->dRecv, (size_t)mc->nLocal * ROW, ncclDouble, mc->comm, mc->stream) != ncclSuccess ||
                cudaStreamSynchronize (mc->stream) != cudaSuccess)
                { mb200_mc3_destroy (mc); return MB200_MC3_ERROR_NCCL; }
            }
        }
    *out = mc;
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_local_chain_count (const mb200_mc3 *mc) { return mc ? mc->nLocal : 0; }
int mb200_mc3_first_local_chain (const mb200_mc3 *mc) { return mc ? mc->first : 0; }

int mb200_mc3_owner (const mb200_mc3 *mc, int g)
{
    if (!mc || g < 0 || g >= mc->nGlobal)
        return MB200_MC3_ERROR_RANGE;
    return OwnerOf (mc, g);
}

int mb200_mc3_chain_id (const mb200_mc3 *mc, int g)
{
    if (!mc || g < 0 || g >= mc->nGlobal)
        return MB200_MC3_ERROR_RANGE;
    return mc->chainId[g];
}

double mb200_mc3_temperature (const mb200_mc3 *mc, int g)
{
    if (!mc || g < 0 || g >= mc->nGlobal)
        return 0.0;
    return TemperatureOf (mc, mc->chainId[g]);
}

double *mb200_mc3_table (mb200_mc3 *mc) { return mc ? mc->table : NULL; }

/* the pairs the coming attempt_swaps will draw, without advancing the generator */
int mb200_mc3_next_swaps_cross_ranks (const mb200_mc3 *mc)
{
    long seed;
    int  run, j, a, b;

    if (!mc || mc->cfg.world == 1 || mc->cfg.chains_per_run < 2)
        return 0;
    seed = mc->swapSeed;
    for (run=0; run<mc->cfg.num_runs; run++)
        for (j=0; j<mc->cfg.num_swaps; j++)
            {
            PickSwappers (mc, &seed, run, &a, &b);
            NextRandom (&seed);                                  /* the acceptance draw */
            if (OwnerOf (mc, a) != OwnerOf (mc, b))
                return 1;
            }
    return 0;
}

int mb200_mc3_exchange_begin (mb200_mc3 *mc, const double *lnl, const double *lnprior)
{
    int i;

    if (!mc || !lnl || !lnprior)
        return MB200_MC3_ERROR_GENERAL;
    if (mc->pending)
        return MB200_MC3_ERROR_PROTOCOL;
    for (i=0; i<mc->nLocal; i++)
        {
        double *row = mc->table + (size_t)(mc->first + i) * ROW;
        row[0] = lnl[i]; row[1] = lnprior[i]; row[2] = (double) mc->chainId[mc->first + i];
        }
    mc->pending = 1;
    mc->pendingCollective = 0;
    if (mc->cfg.world == 1)
        { mc->fresh = 1; return MB200_MC3_SUCCESS; }
    if (mc->cfg.backend == MB200_MC3_LOOPBACK)
        { mc->fresh = 1; return MB200_MC3_SUCCESS; }             /* the caller fills the remote rows */
    if (!mb200_mc3_next_swaps_cross_ranks (mc))
        { mc->fresh = 0; return MB200_MC3_SUCCESS; }             /* every swapper pair is co-resident: no message (src/mcmc.c:668) */
    memcpy (mc->hSend, mc->table + (size_t)mc->first * ROW, (size_t)mc->nLocal * ROW * sizeof(double));
    if (cudaSetDevice (mc->cfg.device) != cudaSuccess ||
        cudaMemcpyAsync (mc->dSend, mc->hSend, (size_t)mc->nLocal * ROW * sizeof(double), cudaMemcpyHostToDevice, mc->stream) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    if (ncclAllGather (mc->dSend, mc->dRecv, (size_t)mc->nLocal * ROW, ncclDouble, mc->comm, mc->stream) != ncclSuccess)
        return MB200_MC3_ERROR_NCCL;
    if (cudaMemcpyAsync (mc->hRecv, mc->dRecv, (size_t)mc->nGlobal * ROW * sizeof(double), cudaMemcpyDeviceToHost, mc->stream) != cudaSuccess ||
        cudaEventRecord (mc->done, mc->stream) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    mc->pendingCollective = 1;
    mc->collectives++;
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_exchange_end (mb200_mc3 *mc)
{
    int g;

    if (!mc)
        return MB200_MC3_ERROR_GENERAL;
    if (!mc->pending)
        return MB200_MC3_ERROR_PROTOCOL;
    mc->pending = 0;
    if (mc->cfg.backend == MB200_MC3_LOOPBACK && mc->cfg.world > 1)
        {
        /* the caller's transport has filled the other processes' rows */
        for (g=0; g<mc->nGlobal; g++)
            if (OwnerOf (mc, g) != mc->cfg.rank)
                mc->chainId[g] = (int) mc->table[(size_t)g*ROW + 2];
        return MB200_MC3_SUCCESS;
        }
    if (!mc->pendingCollective)
        return MB200_MC3_SUCCESS;
    if (cudaEventSynchronize (mc->done) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    memcpy (mc->table, mc->hRecv, (size_t)mc->nGlobal * ROW * sizeof(double));
    for (g=0; g<mc->nGlobal; g++)
        mc->chainId[g] = (int) mc->table[(size_t)g*ROW + 2];     /* the owners' view of the heats */
    mc->fresh = 1;
    mc->pendingCollective = 0;
    return MB200_MC3_SUCCESS;
}

static void HashDecision (unsigned long long *h, int a, int b, int ok)
{
    int v[3], i, j;
    v[0] = a; v[1] = b; v[2] = ok;
    for (i=0; i<3; i++)
        for (j=0; j<4; j++)
            *h = (*h ^ (unsigned long long)((v[i] >> (8*j)) & 0xff)) * 1099511628211ULL;
}

int mb200_mc3_attempt_swaps (mb200_mc3 *mc, int *accepted)
{
    int     run, j, a, b, tmp, nAcc = 0, chI, chJ, n, mine;
    double  tempA, tempB, lnLikeA, lnLikeB, lnPriorA, lnPriorB, lnR, r, u;

    if (!mc)
        return MB200_MC3_ERROR_GENERAL;
    if (mc->pending)
        return MB200_MC3_ERROR_PROTOCOL;
    n = mc->cfg.chains_per_run;
    if (n < 2)
        { if (accepted) *accepted = 0; return MB200_MC3_SUCCESS; }
    for (run=0; run<mc->cfg.num_runs; run++)
        for (j=0; j<mc->cfg.num_swaps; j++)
            {
            PickSwappers (mc, &mc->swapSeed, run, &a, &b);
            u = NextRandom (&mc->swapSeed);
            mine = (OwnerOf (mc, a) == mc->cfg.rank && OwnerOf (mc, b) == mc->cfg.rank);
            if (!mc->fresh && !mine)
                {
                /* both swappers live on another process and nothing was exchanged this generation:
                   their owner decides; this process learns the outcome with the next gathered table */
                mc->runMissed[run]++;
                continue;
                }
            /* AttemptSwap, serial branch (src/mcmc.c:700-735) */
            tempA = TemperatureOf (mc, mc->chainId[a]);
            tempB = TemperatureOf (mc, mc->chainId[b]);
            lnLikeA = mc->table[(size_t)a*ROW];     lnPriorA = mc->table[(size_t)a*ROW + 1];
            lnLikeB = mc->table[(size_t)b*ROW];     lnPriorB = mc->table[(size_t)b*ROW + 1];
            lnR = (tempB * (lnLikeA + lnPriorA) + tempA * (lnLikeB + lnPriorB)) - (tempA * (lnLikeA + lnPriorA) + tempB * (lnLikeB + lnPriorB));
            if (lnR < -100.0)
                r = 0.0;
            else if (lnR > 0.0)
                r = 1.0;
            else
                r = exp (lnR);
            chI = mc->chainId[a]; chJ = mc->chainId[b];
            if (chJ < chI) { tmp = chI; chI = chJ; chJ = tmp; }
            chI %= n; chJ %= n;
            mc->swapInfo[((size_t)run * n + chJ) * n + chI]++;
            if (u < r)
                {
                tmp = mc->chainId[a]; mc->chainId[a] = mc->chainId[b]; mc->chainId[b] = tmp;     /* swap heats, not states */
                mc->table[(size_t)a*ROW + 2] = (double) mc->chainId[a];
                mc->table[(size_t)b*ROW + 2] = (double) mc->chainId[b];
                mc->swapInfo[((size_t)run * n + chI) * n + chJ]++;
                nAcc++;
                }
            HashDecision (&mc->hash, a, b, u < r);
            HashDecision (&mc->runHash[run], a, b, u < r);
            }
    mc->fresh = 0;              /* lnL values age with the next generation */
    if (accepted)
        *accepted = nAcc;
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_reduce_sum (mb200_mc3 *mc, double *values, int count, int root)
{
    if (!mc || !values || count < 1 || root < 0 || root >= mc->cfg.world)
        return MB200_MC3_ERROR_RANGE;
    if (mc->cfg.world == 1 || mc->cfg.backend == MB200_MC3_LOOPBACK)
        return MB200_MC3_SUCCESS;
    if (cudaSetDevice (mc->cfg.device) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    if (count > mc->redCap)
        {
        if (mc->dRed) cudaFree (mc->dRed);
        mc->dRed = NULL; mc->redCap = 0;
        if (cudaMalloc ((void **)&mc->dRed, (size_t)2 * count * sizeof(double)) != cudaSuccess)
            return MB200_MC3_ERROR_CUDA;
        mc->redCap = count;
        }
    if (cudaMemcpyAsync (mc->dRed, values, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, mc->stream) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    if (ncclReduce (mc->dRed, mc->dRed + count, (size_t)count, ncclDouble, ncclSum, root, mc->comm, mc->stream) != ncclSuccess)
        return MB200_MC3_ERROR_NCCL;
    if (mc->cfg.rank == root &&
        cudaMemcpyAsync (values, mc->dRed + count, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, mc->stream) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    if (cudaStreamSynchronize (mc->stream) != cudaSuccess)
        return MB200_MC3_ERROR_CUDA;
    return MB200_MC3_SUCCESS;
}

int mb200_mc3_barrier (mb200_mc3 *mc)
{
    double z = 0.0;
    if (!mc)
        return MB200_MC3_ERROR_GENERAL;
    if (mc->cfg.world == 1 || mc->cfg.backend == MB200_MC3_LOOPBACK)
        return MB200_MC3_SUCCESS;
    return mb200_mc3_reduce_sum (mc, &z, 1, 0);
}

int mb200_mc3_swap_info (const mb200_mc3 *mc, int *out)
{
    if (!mc || !out)
        return MB200_MC3_ERROR_GENERAL;
    memcpy (out, mc->swapInfo, (size_t)mc->cfg.num_runs * mc->cfg.chains_per_run * mc->cfg.chains_per_run * sizeof(int));
    return MB200_MC3_SUCCESS;
}

long long mb200_mc3_collectives (const mb200_mc3 *mc) { return mc ? mc->collectives : 0; }
unsigned long long mb200_mc3_decision_hash (const mb200_mc3 *mc) { return mc ? mc->hash : 0; }

/* per-run view: a process that saw every swap of a run (missed == 0) holds that run's full history */
unsigned long long mb200_mc3_run_hash (const mb200_mc3 *mc, int run)
{
    return (mc && run >= 0 && run < mc->cfg.num_runs) ? mc->runHash[run] : 0;
}

long long mb200_mc3_run_missed (const mb200_mc3 *mc, int run)
{
    return (mc && run >= 0 && run < mc->cfg.num_runs) ? mc->runMissed[run] : -1;
}
