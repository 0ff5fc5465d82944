// mb200_kernels.cuh -- sm_100a kernels of the tree-likelihood hot path.
//
//   tiprobs_kernel      K1  P(t) = max(0, sum_s c_ijs exp(lambda_s t)), double -> float
//                           (TiProbs_Gen, reference src/likelihood.c:9424-9558)
//   eval_nuc4_kernel    K2+K3+K4+K5 fused for S = 4: the whole dirty operation list of an
//                           evaluation (CondLikeDown/Root_NUC4*, CondLikeScaler_NUC4*,
//                           RemoveNodeScalers, Likelihood_NUC4*; src/likelihood.c:786,
//                           1121, 2953, 5137, 5202, 6468, 7981) in ONE launch for ALL
//                           chains of a generation.
//   eval_gen_kernel     same fusion for any S (CondLikeDown/Root_Gen*, CondLikeScaler_Gen*,
//                           Likelihood_Gen*; src/likelihood.c:204, 2152, 4939, 5764)
//
// Why one launch can walk a whole tree: Felsenstein pruning never mixes site patterns, so a
// CTA that owns a tile of patterns can execute every node update of the evaluation for its
// tile, in post-order, without any inter-CTA synchronisation.  A thread re-reads only what it
// wrote itself (same pattern), which CUDA orders without fences.  Grid = (pattern tiles,
// evaluations); the per-node rescaler and the site-scaler add/remove bookkeeping live in
// registers, the root integration and the weighted log-sum close the same kernel, and a
// ticketed last-CTA pass makes the final double sum order-deterministic.
#pragma once
#include "mb200_device.cuh"
#include <cuda_runtime.h>
#include <float.h>

#define MB200_TIME_MIN ((double)1.0E-11f)   /* TIME_MIN is a float literal, src/bayes.h:321 */
#define MB200_TIME_MAX ((double)100.0f)     /* TIME_MAX, src/bayes.h:322                    */
#define MB200_LIKE_EPSILON 1.0e-300         /* src/likelihood.c:44                           */
#define MB200_QUIRK_FLAG 1
#define MB200_SHORTCUT_FLAG 2                /* MB200_FLAG_TIP_SHORTCUTS */
#define MB200_GUARD_FLAG 4                   /* MB200_FLAG_RANGE_GUARD */
#define MB200_GUARD_MIN  1.0e-24f            /* rescaler maxima / unscaled root likelihoods below this trip the guard */
#define MB200_GUARD_LN   (-55.262f)          /* log (MB200_GUARD_MIN) */

#ifdef MB200_PHASE_TIMING
__device__ __forceinline__ unsigned long long mb200_now () { unsigned long long t; asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MB200_STAMP(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (slot) < 64) ctx.dbg[blockIdx.y*64 + (slot)] = mb200_now (); } while (0)
#define MB200_SUBSTAMP(o, sub) do { if ((o) == 3) MB200_STAMP (40 + (sub)); } while (0)
#else
#define MB200_STAMP(slot) do { } while (0)
#define MB200_SUBSTAMP(o, sub) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------
// K1: transition matrices.  grid = (matrix updates, K), block = 128.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tiprobs_kernel (DevCtx ctx, const DevEval *__restrict__ evals, int nEval, const double *__restrict__ dvals,
                const DevMat *__restrict__ mats)
{
    __shared__ double sExp[MB200_DEV_MAX_STATES];
    __shared__ int sEvalIdx;
    const DevMat   mu = mats[blockIdx.x];
    const int      k  = blockIdx.y;
    const int      S  = ctx.S;
    if (threadIdx.x == 0)
        {
        int e = 0;                                // the evaluation whose update list holds this matrix
        while (e + 1 < nEval && (int) blockIdx.x >= evals[e + 1].matOff)
            e++;
        while (e > 0 && evals[e].nMat == 0)       // matOff is non-decreasing; skip empty lists
            e--;
        sEvalIdx = e;
        }
    __syncthreads ();
    const DevEval *ev = evals + sEvalIdx;
    if (ev->fuseP)
        return;                                   // rebuilt inside the pruning kernel
    const double  *rates = dvals + ev->dOff;
    const double  *freqs = rates + 2*ctx.K;
    const double   t  = mu.length * rates[k];
    float         *P  = ctx.matrices + ((size_t)mu.matrix * ctx.K + k) * S * S;

    if (t < MB200_TIME_MIN)
        {
        for (int idx = threadIdx.x; idx < S*S; idx += blockDim.x)
            P[idx] = (idx / S == idx % S) ? 1.0f : 0.0f;
        return;
        }
    if (t > MB200_TIME_MAX)
        {
        for (int idx = threadIdx.x; idx < S*S; idx += blockDim.x)
            P[idx] = (float) freqs[idx % S];
        return;
        }
    const size_t   partLen = 2*(size_t)S + (size_t)S*S*S;
    const double *lam = (mu.eigen == -2) ? (freqs + S)            // eigensystem carried by the evaluation
                                         : ctx.eigen + ((size_t)mu.eigen * ctx.cijkParts + (ctx.cijkParts > 1 ? k : 0)) * partLen;
    const double *cij = lam + 2*S;
    if (threadIdx.x < S)
        sExp[threadIdx.x] = exp (lam[threadIdx.x] * t);
    __syncthreads ();
    for (int idx = threadIdx.x; idx < S*S; idx += blockDim.x)
        {
        const double *c = cij + (size_t)idx * S;
        double sum = 0.0;
        for (int s = 0; s < S; s++)
            sum += c[s] * sExp[s];
        P[idx] = (float) ((sum < 0.0) ? 0.0 : sum);
        }
}

__device__ __forceinline__ float umma_to_tf32 (float x)      // round-to-nearest TF32, as umma::to_tf32
{
    unsigned r;
    asm ("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float (r);
}

// 61-state tensor-core path: entry (i, j) of a P(t) matrix -> the pre-split operand image of tc kernel B
// (canonical K-major layout, umma_common.cuh canon_off; mb200_kernels_tc.cuh)
__device__ __forceinline__ void write_split61 (float *split61, int matrix, int K, int k, int i, int j, float pv)
{
    const float hi = umma_to_tf32 (pv), lo = umma_to_tf32 (pv - hi);
    float *img = split61 + ((size_t)matrix * K + k) * (2 * 64 * 64);
    // one canonical image of 128 rows: rows 0..63 hi, rows 64..127 lo (tc_write_split_entry, mb200_kernels_tc.cuh)
    const unsigned offHi = (unsigned)((j >> 2) * (128 >> 3) * 128 + (i >> 3) * 128 + (i & 7) * 16 + (j & 3) * 4) / 4u;
    img[offHi] = hi;
    img[offHi + (64 >> 3) * 128 / 4] = lo;
}

// K1 for large state counts (61-state codon): the same sum, organised for memory parallelism.
// grid = (matrix updates, K, ceil(S/4)); one warp per ancestral state i: for each j the 32 lanes
// read the S consecutive doubles c[i][j][.] (coalesced), multiply by exp(lambda_s t) from shared
// memory and tree-reduce with shuffles.  (Summation order differs from the reference's sequential
// loop by O(1e-16) relative, invisible after the cast to float.)
__global__ void __launch_bounds__(128)
tiprobs_wide_kernel (DevCtx ctx, const DevEval *__restrict__ evals, int nEval, const double *__restrict__ dvals,
                     const DevMat *__restrict__ mats, float *__restrict__ split61)
{
    __shared__ double sExp[MB200_DEV_MAX_STATES];
    __shared__ int sEvalIdx;
    const DevMat   mu = mats[blockIdx.x];
    const int      k  = blockIdx.y;
    const int      S  = ctx.S;
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        {
        int e = 0;
        while (e + 1 < nEval && (int) blockIdx.x >= evals[e + 1].matOff)
            e++;
        while (e > 0 && evals[e].nMat == 0)
            e--;
        sEvalIdx = e;
        }
    __syncthreads ();
    const DevEval *ev = evals + sEvalIdx;
    if (ev->fuseP)
        return;
    const double  *rates = dvals + ev->dOff;
    const double  *freqs = rates + 2*ctx.K;
    const double   t  = mu.length * rates[k];
    const int      i  = blockIdx.z * 4 + warp;
    float         *P  = ctx.matrices + ((size_t)mu.matrix * ctx.K + k) * S * S;
    if (t < MB200_TIME_MIN || t > MB200_TIME_MAX)
        {
        if (i < S)
            for (int j = lane; j < S; j += 32)
                {
                const float pv = (t < MB200_TIME_MIN) ? ((i == j) ? 1.0f : 0.0f) : (float) freqs[j];
                P[i*S + j] = pv;
                if (split61 != nullptr)
                    write_split61 (split61, mu.matrix, ctx.K, k, i, j, pv);
                }
        return;
        }
    const size_t   partLen = 2*(size_t)S + (size_t)S*S*S;
    const double *lam = ctx.eigen + ((size_t)mu.eigen * ctx.cijkParts + (ctx.cijkParts > 1 ? k : 0)) * partLen;
    const double *cij = lam + 2*S;
    if (threadIdx.x < S)
        sExp[threadIdx.x] = exp (lam[threadIdx.x] * t);
    __syncthreads ();
    if (i >= S)
        return;
    const double e0 = (lane < S) ? sExp[lane] : 0.0, e1 = (lane + 32 < S) ? sExp[lane + 32] : 0.0;
    // four j at a time: the loads of a group are all in flight before the first reduction (one L2 round
    // trip per group instead of one per j)
    for (int j0 = 0; j0 < S; j0 += 4)
        {
        double sum[4];
        #pragma unroll
        for (int u = 0; u < 4; u++)
            {
            const int j = (j0 + u < S) ? j0 + u : S - 1;
            const double *c = cij + ((size_t)i * S + j) * S;
            double v = 0.0;
            if (lane < S)      v  = c[lane] * e0;
            if (lane + 32 < S) v += c[lane + 32] * e1;
            sum[u] = v;
            }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            {
            #pragma unroll
            for (int u = 0; u < 4; u++)
                sum[u] += __shfl_xor_sync (0xffffffffu, sum[u], off);
            }
        if (lane < 4 && j0 + lane < S)
            {
            const double sj = (lane == 0) ? sum[0] : (lane == 1) ? sum[1] : (lane == 2) ? sum[2] : sum[3];
            const float  pv = (float) ((sj < 0.0) ? 0.0 : sj);
            P[i*S + j0 + lane] = pv;
            if (split61 != nullptr)              // the tensor-core kernel's operand image, no extra kernel
                write_split61 (split61, mu.matrix, ctx.K, k, i, j0 + lane, pv);
            }
        }
}

// ---------------------------------------------------------------------------------------
// K1 for large state counts as a batched contraction.  MrBayes hands over c[i][j][s] = V[i][s] * Vinv[s][j]
// (CalcCijk, src/utils.c:9734-9746): for every s the S x S slice is a rank-one matrix, so it factors again into a
// column u_s and a row w_s (any scaling: u_s[i] = c[i][j0][s], w_s[j] = c[i0][j][s] / c[i0][j0][s] with (i0, j0)
// the slice's largest entry).  With U[i][s], W[s][j] in hand
//     P_k = (U diag(e^{lambda t r_k})) W                                   (TiProbs_Gen, src/likelihood.c:9499-9542)
// is an S x S x S matrix product per branch and category: 2 S^2 doubles of operands instead of the S^3 doubles of
// the c_ijk block (1.8 MB at S = 61), both staged in shared memory, 4 x 4 outputs per thread.
// Rounding: (U e) W instead of (U W) e, same summation order over s: <= 1 ulp of double per term, i.e. the
// float-rounded P(t) agrees with the reference's except where the double sum sits on a float rounding boundary.
// ---------------------------------------------------------------------------------------
// grid = (S, eigen parts), block = 256: factor one slice
__global__ void __launch_bounds__(256)
cijk_factor_kernel (const double *__restrict__ block, double *__restrict__ factor, int S)
{
    __shared__ double sMax[256];
    __shared__ int    sArg[256];
    const int s = blockIdx.x, part = blockIdx.y;
    const size_t partLen = 2*(size_t)S + (size_t)S*S*S;
    const double *c = block + (size_t)part * partLen + 2*S;
    double *U = factor + (size_t)part * 2 * S * S, *W = U + (size_t)S * S;
    double best = -1.0; int arg = 0;
    for (int e = threadIdx.x; e < S*S; e += 256)
        {
        const double a = fabs (c[(size_t)e * S + s]);
        if (a > best) { best = a; arg = e; }
        }
    sMax[threadIdx.x] = best; sArg[threadIdx.x] = arg;
    __syncthreads ();
    for (int off = 128; off > 0; off >>= 1)
        {
        if (threadIdx.x < off && (sMax[threadIdx.x + off] > sMax[threadIdx.x] ||
                                  (sMax[threadIdx.x + off] == sMax[threadIdx.x] && sArg[threadIdx.x + off] < sArg[threadIdx.x])))
            { sMax[threadIdx.x] = sMax[threadIdx.x + off]; sArg[threadIdx.x] = sArg[threadIdx.x + off]; }
        __syncthreads ();
        }
    const int i0 = sArg[0] / S, j0 = sArg[0] % S;
    const double piv = c[((size_t)i0 * S + j0) * S + s];
    for (int e = threadIdx.x; e < S; e += 256)
        {
        U[(size_t)e * S + s] = c[((size_t)e * S + j0) * S + s];                                  // u_s[i]
        W[(size_t)s * S + e] = (piv != 0.0) ? c[((size_t)i0 * S + e) * S + s] / piv : 0.0;      // w_s[j]
        }
}

// grid = (matrix updates, K), block = 256, dynamic shared memory = 2 * S * LD doubles (LD = S rounded up to 4)
__global__ void __launch_bounds__(256)
tiprobs_mm_kernel (DevCtx ctx, const DevEval *__restrict__ evals, int nEval, const double *__restrict__ dvals,
                   const DevMat *__restrict__ mats, const double *__restrict__ factor, float *__restrict__ split61)
{
    extern __shared__ __align__(16) double mmS[];
    __shared__ double sExp[MB200_DEV_MAX_STATES];
    __shared__ int sEvalIdx;
    const DevMat   mu = mats[blockIdx.x];
    const int      k  = blockIdx.y;
    const int      S  = ctx.S, LD = (S + 3) & ~3;
    double *sUt = mmS;                      // [s][i]  (U e, transposed: a thread's four rows are contiguous)
    double *sW  = mmS + (size_t)S * LD;     // [s][j]
    if (threadIdx.x == 0)
        {
        int e = 0;
        while (e + 1 < nEval && (int) blockIdx.x >= evals[e + 1].matOff)
            e++;
        while (e > 0 && evals[e].nMat == 0)
            e--;
        sEvalIdx = e;
        }
    __syncthreads ();
    const DevEval *ev = evals + sEvalIdx;
    if (ev->fuseP)
        return;
    const double  *rates = dvals + ev->dOff;
    const double  *freqs = rates + 2*ctx.K;
    const double   t  = mu.length * rates[k];
    float         *P  = ctx.matrices + ((size_t)mu.matrix * ctx.K + k) * S * S;
    if (t < MB200_TIME_MIN || t > MB200_TIME_MAX)
        {
        for (int idx = threadIdx.x; idx < S*S; idx += blockDim.x)
            {
            const int i = idx / S, j = idx % S;
            const float pv = (t < MB200_TIME_MIN) ? ((i == j) ? 1.0f : 0.0f) : (float) freqs[j];
            P[idx] = pv;
            if (split61 != nullptr)
                write_split61 (split61, mu.matrix, ctx.K, k, i, j, pv);
            }
        return;
        }
    const int     part = (ctx.cijkParts > 1) ? k : 0;
    const size_t  partLen = 2*(size_t)S + (size_t)S*S*S;
    const double *lam = ctx.eigen + ((size_t)mu.eigen * ctx.cijkParts + part) * partLen;
    const double *U = factor + ((size_t)mu.eigen * ctx.cijkParts + part) * 2 * S * S, *W = U + (size_t)S * S;
    if (threadIdx.x < S)
        sExp[threadIdx.x] = exp (lam[threadIdx.x] * t);
    for (int idx = threadIdx.x; idx < S * LD; idx += blockDim.x)
        { sUt[idx] = 0.0; sW[idx] = 0.0; }
    __syncthreads ();
    for (int idx = threadIdx.x; idx < S*S; idx += blockDim.x)
        {
        const int a = idx / S, b = idx % S;
        sUt[(size_t)b * LD + a] = U[idx] * sExp[b];          // U[i = a][s = b] e_s  ->  [s][i]
        sW [(size_t)a * LD + b] = W[idx];                    // W[s = a][j = b]
        }
    __syncthreads ();
    const int nt = LD / 4;                                   // 4 x 4 micro-tiles per side
    for (int tile = threadIdx.x; tile < nt * nt; tile += blockDim.x)
        {
        const int i4 = (tile / nt) * 4, j4 = (tile % nt) * 4;
        double acc[4][4];
        #pragma unroll
        for (int a = 0; a < 4; a++)
            #pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
        for (int s = 0; s < S; s++)
            {
            const double2 u01 = *reinterpret_cast<const double2 *>(sUt + (size_t)s * LD + i4), u23 = *reinterpret_cast<const double2 *>(sUt + (size_t)s * LD + i4 + 2);
            const double2 w01 = *reinterpret_cast<const double2 *>(sW + (size_t)s * LD + j4),  w23 = *reinterpret_cast<const double2 *>(sW + (size_t)s * LD + j4 + 2);
            const double u[4] = { u01.x, u01.y, u23.x, u23.y }, w[4] = { w01.x, w01.y, w23.x, w23.y };
            #pragma unroll
            for (int a = 0; a < 4; a++)
                #pragma unroll
                for (int b = 0; b < 4; b++)
                    acc[a][b] = fma (u[a], w[b], acc[a][b]);
            }
        #pragma unroll
        for (int a = 0; a < 4; a++)
            #pragma unroll
            for (int b = 0; b < 4; b++)
                if (i4 + a < S && j4 + b < S)
                    {
                    const float pv = (float) ((acc[a][b] < 0.0) ? 0.0 : acc[a][b]);
                    P[(i4 + a) * S + j4 + b] = pv;
                    if (split61 != nullptr)
                        write_split61 (split61, mu.matrix, ctx.K, k, i4 + a, j4 + b, pv);
                    }
        }
}

// c_ijk = V[i][k] * Vinv[k][j]  (CalcCijk, src/utils.c:9734-9746)
__global__ void cijk_kernel (double *block, const double *V, const double *Vinv, const double *lambda, int S)
{
    const size_t n3 = (size_t)S*S*S;
    for (size_t idx = blockIdx.x*(size_t)blockDim.x + threadIdx.x; idx < n3; idx += (size_t)gridDim.x*blockDim.x)
        {
        int k = (int)(idx % S);
        int j = (int)((idx / S) % S);
        int i = (int)(idx / ((size_t)S*S));
        block[2*S + idx] = V[i*S + k] * Vinv[k*S + j];
        }
    if (blockIdx.x == 0)
        for (int s = threadIdx.x; s < S; s += blockDim.x)
            {
            block[s]     = lambda[s];
            block[S + s] = 0.0;
            }
}

// invMask[c] = AND over tips of tip64[tip][c]
__global__ void invmask_kernel (uint64_t *inv, const uint64_t *tip64, int tipCount, int C)
{
    int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= C) return;
    uint64_t m = ~(uint64_t)0;
    for (int t = 0; t < tipCount; t++)
        m &= tip64[(size_t)t*C + c];
    inv[c] = m;
}

// ---------------------------------------------------------------------------------------
// deterministic block reduction of (double term, int abort) + ticketed cross-tile sum
// ---------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void finish_lnl (const DevCtx &ctx, int evalIdx, double term, int abortFlag,
                                            DevResult *out, int seq)
{
    __shared__ double sSum[NT/32];
    __shared__ int    sAb[NT/32];
    __shared__ int    sLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    #pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        {
        term      += __shfl_xor_sync (0xffffffffu, term, off);
        abortFlag |= __shfl_xor_sync (0xffffffffu, abortFlag, off);
        }
    if (lane == 0) { sSum[warp] = term; sAb[warp] = abortFlag; }
    __syncthreads ();
    if (threadIdx.x == 0)
        {
        double s = 0.0; int a = 0;
        #pragma unroll
        for (int w = 0; w < NT/32; w++) { s += sSum[w]; a |= sAb[w]; }
        if (ctx.hostSum)
            {
            // latency path: no ticket, no second pass -- the tile's partial goes straight to the
            // caller (one 16-byte store into mapped host memory), which sums the few tiles itself
            int4 pkt;
            pkt.x = __double2loint (s); pkt.y = __double2hiint (s); pkt.z = a; pkt.w = seq;
            *reinterpret_cast<int4 *>(&out[(size_t)evalIdx*ctx.numTiles + blockIdx.x]) = pkt;
            sLast = 0;
            }
        else
            {
        ctx.tilePartial[(size_t)evalIdx*ctx.numTiles + blockIdx.x] = s;
        ctx.tileAbort  [(size_t)evalIdx*ctx.numTiles + blockIdx.x] = a;
        __threadfence ();
        unsigned int t = atomicAdd (&ctx.ticket[evalIdx], 1u);
        sLast = (t == (unsigned int)ctx.numTiles - 1u);
            }
        }
    __syncthreads ();
    if (!sLast)
        return;
    __threadfence ();
    // last CTA of this evaluation: fixed-order sum over the tiles
    if (ctx.numTiles <= MB200_SEQ_SUM_TILES)
        {
        // few tiles: plain left-to-right sum, the order the host uses in hostSum mode
        if (threadIdx.x == 0)
            {
            double tot = 0.0; int ab = 0;
            for (int tIdx = 0; tIdx < ctx.numTiles; tIdx++)
                {
                tot += __ldcg (&ctx.tilePartial[(size_t)evalIdx*ctx.numTiles + tIdx]);
                ab  |= __ldcg (&ctx.tileAbort  [(size_t)evalIdx*ctx.numTiles + tIdx]);
                }
            const double lnL = ab ? -DBL_MAX : tot;
            int4 pkt;
            pkt.x = __double2loint (lnL); pkt.y = __double2hiint (lnL); pkt.z = ab ? 1 : 0; pkt.w = seq;
            *reinterpret_cast<int4 *>(&out[evalIdx]) = pkt;
            ctx.ticket[evalIdx] = 0u;
            }
        return;
        }
    double s = 0.0; int a = 0;
    for (int tIdx = threadIdx.x; tIdx < ctx.numTiles; tIdx += NT)
        {
        s += __ldcg (&ctx.tilePartial[(size_t)evalIdx*ctx.numTiles + tIdx]);
        a |= __ldcg (&ctx.tileAbort  [(size_t)evalIdx*ctx.numTiles + tIdx]);
        }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        {
        s += __shfl_xor_sync (0xffffffffu, s, off);
        a |= __shfl_xor_sync (0xffffffffu, a, off);
        }
    __syncthreads ();
    if (lane == 0) { sSum[warp] = s; sAb[warp] = a; }
    __syncthreads ();
    if (threadIdx.x == 0)
        {
        double tot = 0.0; int ab = 0;
        #pragma unroll
        for (int w = 0; w < NT/32; w++) { tot += sSum[w]; ab |= sAb[w]; }
        // one 16-byte store: lnL, status and the sequence number travel in a single write, so a
        // host polling `seq` in (mapped, pinned) memory never sees a half-written result and no
        // system-scope fence sits on the critical path
        const double lnL = ab ? -DBL_MAX : tot;
        int4 pkt;
        pkt.x = __double2loint (lnL); pkt.y = __double2hiint (lnL); pkt.z = ab ? 1 : 0; pkt.w = seq;
        *reinterpret_cast<int4 *>(&out[evalIdx]) = pkt;
        ctx.ticket[evalIdx] = 0u;            // ready for the next launch
        }
}

// site-likelihood -> weighted log term, with the invariable-sites mixing rules of
// Likelihood_NUC4_* (quirk) and Likelihood_Gen* (src/likelihood.c:5836-5912, 6573-6625)
__device__ __forceinline__ double site_term (double like, double likeI, int hasPInvar, int quirk,
                                             float lnScaler, float weight, int &abortFlag)
{
    double lnLike;
    if (!hasPInvar)
        {
        if (like < MB200_LIKE_EPSILON) { abortFlag = 1; return 0.0; }
        lnLike = (double)lnScaler + log (like);
        }
    else if (quirk)
        {
        if (lnScaler < -200.0f)
            {
            if (likeI > 1E-70)
                like = likeI;
            }
        else
            like = like + (likeI / exp ((double)lnScaler));
        if (like < MB200_LIKE_EPSILON) { abortFlag = 1; return 0.0; }
        lnLike = (double)lnScaler + log (like);
        }
    else
        {
        if (lnScaler < -200.0f)
            {
            if (likeI > 1E-70)
                lnLike = log (likeI);
            else
                lnLike = log (like) + (double)lnScaler;
            }
        else
            lnLike = log (like + (likeI / exp ((double)lnScaler))) + (double)lnScaler;
        if (like < MB200_LIKE_EPSILON) { abortFlag = 1; return 0.0; }
        }
    return lnLike * (double)weight;
}

// ---------------------------------------------------------------------------------------
// S = 4 fused evaluation.  grid = (tiles of NT patterns, evaluations), block = NT.
// One thread owns one site pattern: all K rate categories x 4 states = K float4 registers,
// so the rescaler (max over k and states) needs no shuffle at all; warp shuffles are used
// only in the final lnL reduction.  Every global access is a fully coalesced 16-byte (CL),
// 4-byte (scalers, weights) or 1-byte (tip codes) per-thread access.
// ---------------------------------------------------------------------------------------
// (float) log ((double) m) for the node scalers (CondLikeScaler_NUC4 / _SSE, src/likelihood.c:5183, 5328): the
// library's double-precision log costs ~75 instructions, most of them for arguments that cannot occur
// here.  m is a positive normal float (a rescaler maximum): m = 2^e f with f in [0.7071, 1.4142],
// log f = 2 atanh s, s = (f-1)/(f+1) (|s| <= 0.1716), evaluated in double to ~1e-15 -- the float cast
// then agrees with the library's in all but ~1e-7 of the arguments.  Anything else goes to the library.
__device__ __forceinline__ float log_of_max (float m)
{
    const int bits = __float_as_int (m);
    if (bits < 0x00800000 || bits >= 0x7f800000)
        return (float) log ((double) m);
    int   e = (bits >> 23) - 127;
    float f = __int_as_float ((bits & 0x007fffff) | 0x3f800000);           // [1, 2)
    if (f > 1.41421354f) { f *= 0.5f; e += 1; }
    const double fd = (double) f, dp1 = fd + 1.0, dm1 = fd - 1.0;
    float r0;
    asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(f + 1.0f));
    double r = (double) r0;
    r = fma (r, fma (-dp1, r, 1.0), r);                                    // 1 / (f + 1)
    double sq = dm1 * r;
    sq = fma (fma (-sq, dp1, dm1), r, sq);                                 // s = (f - 1) / (f + 1)
    const double s2 = sq * sq;
    double p = 1.0 / 19.0;
    p = fma (p, s2, 1.0 / 17.0); p = fma (p, s2, 1.0 / 15.0); p = fma (p, s2, 1.0 / 13.0);
    p = fma (p, s2, 1.0 / 11.0); p = fma (p, s2, 1.0 / 9.0);  p = fma (p, s2, 1.0 / 7.0);
    p = fma (p, s2, 1.0 / 5.0);  p = fma (p, s2, 1.0 / 3.0);
    const double two_s = sq + sq;
    const double lf = fma (two_s, p * s2, two_s);                          // 2 atanh s
    return (float) fma ((double) e, 0.69314718055994530942, lf);
}

// r / m for the four states of a rescaled vector (CondLikeScaler_NUC4, src/likelihood.c:5169-5200):
// one reciprocal refined to < 1 ulp, then per element the quotient with one exact-remainder
// correction -- the correctly rounded quotient an IEEE divide returns, at a third of the
// instructions of four divisions.  Outside the exponent range where that argument holds: plain '/'.
__device__ __forceinline__ void scale4 (float4 &r, float m)
{
    if (m > 1.0e-30f && m < 1.0e30f)
        {
        float rc;
        asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(m));
        rc = fmaf (rc, fmaf (-m, rc, 1.0f), rc);
        float q, rem;
        q = r.x * rc; rem = fmaf (-m, q, r.x); r.x = fmaf (rem, rc, q);
        q = r.y * rc; rem = fmaf (-m, q, r.y); r.y = fmaf (rem, rc, q);
        q = r.z * rc; rem = fmaf (-m, q, r.z); r.z = fmaf (rem, rc, q);
        q = r.w * rc; rem = fmaf (-m, q, r.w); r.w = fmaf (rem, rc, q);
        }
    else if (m == 0.0f)
        {
        // every state of every category is zero (a dead pattern): the reference's 0/0.  No reason to
        // spend four slow-path divisions on it
        const float q = __int_as_float (0x7fc00000);
        r.x = (r.x == 0.0f) ? q : r.x / m; r.y = (r.y == 0.0f) ? q : r.y / m;
        r.z = (r.z == 0.0f) ? q : r.z / m; r.w = (r.w == 0.0f) ? q : r.w / m;
        }
    else
        { r.x /= m; r.y /= m; r.z /= m; r.w /= m; }
}

__device__ __forceinline__ float dot4_fma (const float4 p, const float4 x)
{
    // same operation order as CondLikeDown_NUC4_FMA (src/likelihood.c:1149-1169)
    return fmaf (p.w, x.w, fmaf (p.z, x.z, fmaf (p.y, x.y, p.x * x.x)));
}

__device__ __forceinline__ float4 matvec4 (const float4 *rows, const float4 x)
{
    return make_float4 (dot4_fma (rows[0], x), dot4_fma (rows[1], x), dot4_fma (rows[2], x), dot4_fma (rows[3], x));
}

// the same with the four rows of P(t) read from shared memory at a 32-bit shared address
// (rows 16 bytes apart): keeps one live register per base instead of a generic pointer
__device__ __forceinline__ float4 lds128 (unsigned saddr)
{
    float4 v;
    asm volatile ("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

__device__ __forceinline__ float4 matvec4s (unsigned saddr, const float4 x)
{
    const float4 r0 = lds128 (saddr), r1 = lds128 (saddr + 16), r2 = lds128 (saddr + 32), r3 = lds128 (saddr + 48);
    return make_float4 (dot4_fma (r0, x), dot4_fma (r1, x), dot4_fma (r2, x), dot4_fma (r3, x));
}

// contribution of a tip child: sum of the P columns its state set selects == the dense 0/1
// matvec of the reference, bit for bit (adding exact zeros changes nothing)
__device__ __forceinline__ float tip_dot4 (const float4 p, int mask)
{
    float r = (mask & 1) ? p.x : 0.0f;
    if (mask & 2) r += p.y;
    if (mask & 4) r += p.z;
    if (mask & 8) r += p.w;
    return r;
}

// ---------------------------------------------------------------------------------------
// S = 4: fused evaluation (K2 + K3 + K4 + K5, and K1 when FUSE).
//
// Thread mapping.  One site pattern is spread over L = pow2ceil(K) adjacent lanes, one rate
// category each: a lane owns one float4 (the 4 states) per conditional-likelihood vector.  The
// rescaler's max over categories is a log2(L)-step xor-shuffle; everything else is lane-local.
//
// The operation list of an evaluation is cut (on the host) into chunks of <= OPC nodes that touch
// <= MAXS distinct branches.  Per chunk the CTA
//   1. stages the chunk's node list and branch list into shared memory;
//   2. fills one shared-memory slot per branch with the K x 4 rows of P(t): rebuilt in double
//      precision from the eigensystem when the branch is dirty and FUSE is on (TiProbs_Gen; the CTA
//      of tile 0 also publishes it to the matrix buffer), copied from the matrix buffer otherwise;
//   3. walks the nodes WITHOUT any barrier -- a thread only ever touches its own pattern: child
//      loads are prefetched one node ahead, the child that is the previous node's result stays in
//      registers, tips are 1-byte state masks expanded to 0/1 vectors in registers, so every child
//      takes the same matvec path; the scaler maxima go to shared memory;
//   4. takes the logarithms of the chunk's scalers in one batched pass (full ILP instead of a
//      ~50-instruction dependent chain per node), writes the node scalers, reads the old ones, and
//      replays the site-scaler additions in the reference's order (... - old(o) + new(o) ...), so
//      the float rounding sequence is the reference's.
// Root integration, the weighted log-sum and the ticketed cross-tile reduction close the kernel.
//
// FUSE = true : small launches (latency-bound: one warp per scheduler, the dependent instruction
//               chain per node is what counts).  No separate P(t) kernel, no launch gap.
// FUSE = false: large grids (issue/bandwidth-bound).  P(t) comes from tiprobs_kernel once instead
//               of once per CTA; no double-precision exp code, so more CTAs fit per SM.
// ---------------------------------------------------------------------------------------
template <int K> struct Nuc4Geom
{
    static constexpr int L = (K <= 1) ? 1 : (K <= 2) ? 2 : (K <= 4) ? 4 : 8;   // lanes per pattern
    static constexpr int MAXS = (256 / K > 96) ? 96 : 256 / K;                 // P(t) slots per chunk
};

// shared memory of the 4-state kernel (dynamic: more than the 48 KB a static allocation may take)
template <int K, int NT, bool FUSE> struct Nuc4Smem
{
    static constexpr int L    = Nuc4Geom<K>::L;
    static constexpr int MAXS = nuc_maxs (K, FUSE);
    static constexpr int PPB  = NT / L;
    static constexpr int OPC  = nuc_opc (PPB, FUSE);
    static constexpr int MAXT = nuc_maxt (K, FUSE);
    float4 sP[MAXS][K][5];                       // P(t) rows of every branch the chunk touches (4 rows + 1 pad: bank spread)
    float4 sTab[MAXT][16][K];                    // per tip operand, state mask and category: sum of the P(t) columns the mask selects
                                                 // (mask-major: the K lanes of a pattern read one contiguous 16K-byte line)
    double sExp[FUSE ? MAXS : 1][K][4];          // exp(lambda_s t) of the dirty branches
    double sD[2*K + 4];                          // rates[K], catW[K], freqs[4]
    double sEig[72];                               // lambda_re[4], lambda_im[4], c_ijk[64] of slot eigen0
    DevMat sMat[MAXS];
    DevOp  sOps[OPC];
    DevEval sEv;
    DevChunk sCh;
    uint2  sTipInfo[MAXT];                       // .x: byte index of the tip row (| 1u<<31: shortcut applies), .y: P(t) slot offset
    float  sNew[OPC][PPB];                       // per node: scaler maximum, later its logarithm
    float  sOld[OPC][PPB];                       // per node: the old node scaler to remove
    unsigned char sMask[MAXT][PPB];              // the chunk's tip state masks, this CTA's patterns
    unsigned sPreList[NUC_MAXPRE];               // latency path: buffer offsets of the operands fetched at chunk start
    float4 sPre[FUSE ? NUC_MAXPRE : 1][NT];      // ... and their vectors, one per thread (thread-private: no barrier needed)
};

template <int K, int NT, bool FUSE>
__device__ __forceinline__ void
nuc4_body (const DevCtx &ctx, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
           const DevChunk *__restrict__ chunks, const DevMat *__restrict__ cmats,
           const DevOp *__restrict__ ops, DevResult *out, int seq, const JobIndex &jx)
{
    constexpr int L    = Nuc4Geom<K>::L;
    constexpr int PPB  = NT / L;                 // patterns per CTA
    extern __shared__ __align__(16) unsigned char nuc_smem[];
    Nuc4Smem<K, NT, FUSE> &sm = *reinterpret_cast<Nuc4Smem<K, NT, FUSE> *>(nuc_smem);
    auto &sP = sm.sP;   auto &sExp = sm.sExp; auto &sTab = sm.sTab; auto &sMat = sm.sMat; auto &sOps = sm.sOps;
    auto &sNew = sm.sNew; auto &sOld = sm.sOld; auto &sEv = sm.sEv; auto &sCh = sm.sCh; auto &sD = sm.sD;
    auto &sEig = sm.sEig; auto &sTipInfo = sm.sTipInfo; auto &sMask = sm.sMask; auto &sPreList = sm.sPreList; auto &sPre = sm.sPre;

    MB200_STAMP (0);
    // ---- 0. staging of the evaluation and of its first chunk.  With a job index (small launches) all
    //      of it is one round of independent loads; otherwise the header comes first ----
    const bool indexed = (int) blockIdx.y < jx.n;
    DevChunk ch0;
    int dOff0, eig0;
    if (indexed)
        {
        const JobIndexEntry je = jx.e[blockIdx.y];
        ch0.opOff = je.opOff; ch0.nOp = je.nOp; ch0.matOff = je.matOff; ch0.nMat = je.nMat;
        dOff0 = je.dOff; eig0 = je.eigen0;
        }
    if (!indexed)
        {
        if (threadIdx.x < (int)(sizeof(DevEval) / 4))
            reinterpret_cast<int *>(&sEv)[threadIdx.x] = reinterpret_cast<const int *>(evals + blockIdx.y)[threadIdx.x];
        __syncthreads ();
        ch0 = sEv.chunk0; dOff0 = sEv.dOff; eig0 = sEv.eigen0;
        }
    MB200_STAMP (1);
    // chunk lists -> shared memory.  All loads of a thread are issued before its first store (a load
    // followed by its store, loop after loop, would serialise one cold miss per list)
    auto stageChunk = [&] (const DevChunk &ch, bool withEval)
        {
        const int nMatW = (ch.nMat & 0xffff) * 4, nOpW = ch.nOp * (int)(sizeof(DevOp)/4);
        const int *srcE = reinterpret_cast<const int *>(evals + blockIdx.y);
        const int *srcM = reinterpret_cast<const int *>(cmats + ch.matOff);
        const int *srcO = reinterpret_cast<const int *>(ops + ch.opOff);
        int vE = 0, vM = 0, vO = 0, vO2 = 0;
        double vD = 0.0, vG = 0.0;
        const int t = threadIdx.x;
        if (withEval)
            {
            if (indexed && t < (int)(sizeof(DevEval) / 4)) vE = srcE[t];
            if (t < 2*K + 4) vD = dvals[dOff0 + t];
            if (FUSE && t < 72)
                vG = (eig0 == -2) ? dvals[dOff0 + 2*K + 4 + t]          // eigensystem carried by the evaluation
                                  : ctx.eigen[(size_t)eig0 * 72 + t];
            }
        if (t < nMatW)      vM  = srcM[t];
        if (t < nOpW)       vO  = srcO[t];
        if (t + NT < nOpW)  vO2 = srcO[t + NT];
        if (withEval)
            {
            if (indexed && t < (int)(sizeof(DevEval) / 4)) reinterpret_cast<int *>(&sEv)[t] = vE;
            if (t < 2*K + 4) sD[t] = vD;
            if (FUSE && t < 72) sEig[t] = vG;
            }
        if (t < nMatW)      reinterpret_cast<int *>(sMat)[t] = vM;
        if (t < nOpW)       reinterpret_cast<int *>(sOps)[t] = vO;
        if (t + NT < nOpW)  reinterpret_cast<int *>(sOps)[t + NT] = vO2;
        for (int e = t + NT; e < nMatW; e += NT)      reinterpret_cast<int *>(sMat)[e] = srcM[e];
        for (int e = t + 2 * NT; e < nOpW; e += NT)   reinterpret_cast<int *>(sOps)[e] = srcO[e];
        };
    // tip operand list of the chunk, from the staged node list
    auto listTips = [&] (int nOp)
        {
        if (threadIdx.x < nOp)
            {
            const NucOp &o = reinterpret_cast<const NucOp *>(sOps)[threadIdx.x];
            const unsigned kinds = o.kinds;
            #pragma unroll
            for (int j = 0; j < 3; j++)
                {
                const unsigned kind = (kinds >> (4*j)) & 15u;
                if (kind & NUC_TIP)
                    sTipInfo[(kinds >> (13 + 6*j)) & 63u] =
                        make_uint2 (((j == 0) ? o.a1 : (j == 1) ? o.a2 : o.a3) | ((kind == NUC_TIP_ONE) ? 0x80000000u : 0u),
                                    (j == 0) ? o.sp1 : (j == 1) ? o.sp2 : o.sp3);
                if (FUSE && kind == NUC_PRE)
                    sPreList[((unsigned) o.pad >> (4*j)) & 15u] = (j == 0) ? o.a1 : (j == 1) ? o.a2 : o.a3;
                }
            }
        };
    stageChunk (ch0, true);
    __syncthreads ();
    listTips (ch0.nOp);
    __syncthreads ();
    const int   C      = ctx.C;
    const int   lk     = threadIdx.x % L;                   // this lane's rate category
    const int   kk     = (lk < K) ? lk : K - 1;
    const int   pl     = threadIdx.x / L;                   // pattern slot within the CTA
    // the pattern tile this CTA works on; in throughput mode (ctx.patternTiles > gridDim.x, single-chunk
    // evaluations only) a CTA walks several tiles, reusing the P(t) slots and tip tables it built
    int   c0     = blockIdx.x * PPB;
    int   c      = c0 + pl;
    bool  active = (c < C) && (lk < K);
    int   cc     = (c < C) ? c : C - 1;
    float4 *partials4 = reinterpret_cast<float4 *>(ctx.partials);
    const unsigned groupBase = (threadIdx.x & 31) & ~(L - 1);
    const int   nChunk = sEv.nChunk;
    float  lnScaler = (sEv.siteSrc >= 0) ? ctx.scalers[(size_t)sEv.siteSrc * C + cc] : 0.0f;

    float4 cur = make_float4 (0.f, 0.f, 0.f, 0.f);

    // per-thread addressing: everything in the node loop is  base + (uniform offset from the op record)
    // (32-bit element offsets: pack() guarantees they fit; one live register per base)
    unsigned             tOff  = (unsigned) kk * (unsigned) C + (unsigned) cc;
    const unsigned       sPk   = (unsigned) __cvta_generic_to_shared (&sP[0][kk][0]);
    float               *sNewT = &sNew[0][pl];
    const NucOp         *nops  = reinterpret_cast<const NucOp *>(sOps);

    const unsigned sTabK  = (unsigned) __cvta_generic_to_shared (&sTab[0][0][kk]);
    const unsigned sMaskP = (unsigned) __cvta_generic_to_shared (&sMask[0][pl]);

    double termAcc = 0.0; int abortAcc = 0;       // this thread's lnL terms over the tiles of the CTA
    for (int ci = 0; ci < nChunk; ci++)
        {
        // ---- 1. chunk descriptor, node list, branch list (chunk 0: staged above) ----
        if (ci > 0)
            {
            __syncthreads ();                     // previous chunk completely done: shared arrays free
            if (threadIdx.x < 4)
                reinterpret_cast<int *>(&sCh)[threadIdx.x] = reinterpret_cast<const int *>(chunks + sEv.chunkOff + ci - 1)[threadIdx.x];
            __syncthreads ();
            stageChunk (sCh, false);
            __syncthreads ();
            listTips (sCh.nOp);
            __syncthreads ();
            }
        const DevChunk ch = (ci == 0) ? ch0 : sCh;
        const int nMatC = ch.nMat & 0xffff, nTipC = (ch.nMat >> 16) & 0xff, nPreC = FUSE ? (int)((unsigned) ch.nMat >> 24) : 0;
        if (ci == 0) MB200_STAMP (2);

        for (int tIdx = blockIdx.x, firstTile = 1; tIdx < ctx.patternTiles; tIdx += gridDim.x, firstTile = 0)
        {
        if (!firstTile)
            {
            __syncthreads ();                     // the previous tile is done with sMask / sNew / sOld / sPre
            c0 = tIdx * PPB; c = c0 + pl; active = (c < C) && (lk < K); cc = (c < C) ? c : C - 1;
            tOff = (unsigned) kk * (unsigned) C + (unsigned) cc;
            lnScaler = (sEv.siteSrc >= 0) ? ctx.scalers[(size_t)sEv.siteSrc * C + cc] : 0.0f;
            cur = make_float4 (0.f, 0.f, 0.f, 0.f);
            }

        // ---- 2. P(t) slots (K1 fused: TiProbs_Gen, src/likelihood.c:9499-9542); the chunk's tip masks
        //      for this CTA's patterns: the loads go out first, eight deep (the bytes may come from HBM
        //      behind the write stream), and land in shared memory after the P(t) work ----
        constexpr int MKD = 8;
        unsigned char mk[MKD];
        #pragma unroll
        for (int u = 0; u < MKD; u++)
            {
            const int e = threadIdx.x + u * NT;
            if (e < nTipC * PPB)
                {
                const int cp = c0 + (e % PPB);
                mk[u] = ctx.tip8[(sTipInfo[e / PPB].x & 0x7fffffffu) + (unsigned)((cp < C) ? cp : C - 1)];
                }
            }
        // latency path: interior operands from buffers this evaluation does not write, requested now
        float4 pre[NUC_MAXPRE];
        if (FUSE)
            {
            #pragma unroll
            for (int u = 0; u < NUC_MAXPRE; u++)
                if (u < nPreC)
                    pre[u] = partials4[tOff + sPreList[u]];
            }
        for (int r = threadIdx.x; firstTile && r < nMatC * K * 4; r += NT)
            {
            const int s = r & 3, k = (r >> 2) % K, m = r / (4*K);
            const int eg = sMat[m].eigen;
            if (FUSE && eg != -1)
                {
                const double lam = (eg == eig0) ? sEig[s] : ctx.eigen[(size_t)eg * 72 + s];
                sExp[m][k][s] = exp (lam * (sMat[m].length * sD[k]));
                }
            else                                  // clean branch (or P(t) prepared by tiprobs_kernel): copy row s
                sP[m][k][s] = reinterpret_cast<const float4 *>(ctx.matrices + (size_t)sMat[m].matrix * K * 16)[k*4 + s];
            }
        #pragma unroll
        for (int u = 0; u < MKD; u++)
            {
            const int e = threadIdx.x + u * NT;
            if (e < nTipC * PPB)
                sMask[e / PPB][e % PPB] = mk[u];
            }
        if (FUSE)
            {
            #pragma unroll
            for (int u = 0; u < NUC_MAXPRE; u++)
                if (u < nPreC)
                    sPre[u][threadIdx.x] = pre[u];     // thread-private slot: read back by this thread only
            }
        for (int e = threadIdx.x + MKD * NT; e < nTipC * PPB; e += NT)      // more than eight per thread: K < 4 only
            {
            const int cp = c0 + (e % PPB);
            sMask[e / PPB][e % PPB] = ctx.tip8[(sTipInfo[e / PPB].x & 0x7fffffffu) + (unsigned)((cp < C) ? cp : C - 1)];
            }
        __syncthreads ();
        if (ci == 0) MB200_STAMP (50);
        if (FUSE && firstTile)
            {
            for (int r = threadIdx.x; r < nMatC * K * 4; r += NT)
                {
                const int i = r & 3, k = (r >> 2) % K, m = r / (4*K);
                const int eg = sMat[m].eigen;
                if (eg == -1)
                    continue;
                const double t = sMat[m].length * sD[k];
                float4 row;
                if (t < MB200_TIME_MIN)
                    row = make_float4 (i == 0 ? 1.f : 0.f, i == 1 ? 1.f : 0.f, i == 2 ? 1.f : 0.f, i == 3 ? 1.f : 0.f);
                else if (t > MB200_TIME_MAX)
                    row = make_float4 ((float) sD[2*K], (float) sD[2*K+1], (float) sD[2*K+2], (float) sD[2*K+3]);
                else
                    {
                    const double e0 = sExp[m][k][0], e1 = sExp[m][k][1], e2 = sExp[m][k][2], e3 = sExp[m][k][3];
                    float v[4];
                    if (eg == eig0)               // the evaluation's own eigensystem: staged in shared memory
                        {
                        const double *cij = sEig + 8 + i*16;
                        #pragma unroll
                        for (int j = 0; j < 4; j++)
                            {
                            double sum = 0.0;
                            sum += cij[j*4+0] * e0; sum += cij[j*4+1] * e1; sum += cij[j*4+2] * e2; sum += cij[j*4+3] * e3;
                            v[j] = (float) ((sum < 0.0) ? 0.0 : sum);
                            }
                        }
                    else
                        {
                        const double *cij = ctx.eigen + (size_t)eg * 72 + 8 + i*16;
                        #pragma unroll
                        for (int j = 0; j < 4; j++)
                            {
                            double sum = 0.0;
                            sum += cij[j*4+0] * e0; sum += cij[j*4+1] * e1; sum += cij[j*4+2] * e2; sum += cij[j*4+3] * e3;
                            v[j] = (float) ((sum < 0.0) ? 0.0 : sum);
                            }
                        }
                    row = make_float4 (v[0], v[1], v[2], v[3]);
                    }
                sP[m][k][i] = row;
                if (blockIdx.x == 0)              // tile 0 publishes the rebuilt matrices
                    reinterpret_cast<float4 *>(ctx.matrices + (size_t)sMat[m].matrix * K * 16)[k*4 + i] = row;
                }
            __syncthreads ();
            if (ci == 0) MB200_STAMP (51);
            }
        // tip lookup tables: entry[mask][i] = sum over the states j in the mask of P[i][j], added in state
        // order -- the value the reference's dense 0/1 matvec produces (CondLikeDown_NUC4*: products by
        // 0 and 1 are exact, adding 0 changes nothing), so the node loop replaces a tip's matvec by one
        // 16-byte load.  One thread per (tip operand, category, row i): entry[m] = entry[m without its
        // highest state] + P[i][highest state].  Under the scalar kernels' shortcut a missing
        // observation on a tip without partial ambiguity contributes exactly 1.0 (preLike tables,
        // src/likelihood.c:816-832)
        for (int e = threadIdx.x; firstTile && e < nTipC * K * 4; e += NT)
            {
            const int i = e & 3, k = (e >> 2) % K, t = e / (4*K);
            const uint2 ti = sTipInfo[t];
            const float4 p = lds128 ((unsigned) __cvta_generic_to_shared (&sP[0][0][0]) + (unsigned) k * 80u + ti.y + 16u * i);
            float E[16];
            E[0] = 0.0f;
            E[1] = p.x;  E[2] = p.y;  E[4] = p.z;  E[8] = p.w;
            E[3] = E[1] + p.y;
            E[5] = E[1] + p.z;  E[6] = E[2] + p.z;  E[7] = E[3] + p.z;
            #pragma unroll
            for (int m = 1; m < 8; m++)
                E[8 + m] = E[m] + p.w;
            if (ti.x & 0x80000000u)
                E[15] = 1.0f;
            float *dst = reinterpret_cast<float *>(&sTab[t][0][k]) + i;
            #pragma unroll
            for (int m = 0; m < 16; m++)
                dst[m * K * 4] = E[m];
            }
        __syncthreads ();
        if (ci == 0) MB200_STAMP (3);

        // ---- 3. node loop: no barrier, a thread only ever touches its own pattern.  Interior operands
        //      of node n+1 are fetched while node n computes; the two operand sets alternate (xa, xb)
        //      so that no register copies are needed; tip operands are table lookups ----
        const int nOp = ch.nOp;
        float4   xa[3], xb[3];
        // operand j of a node, fetched one node ahead: interior child -> its conditional likelihoods
        // (one 16-byte load); tip child -> its contribution, looked up by state mask
        const unsigned sPreT = (unsigned) __cvta_generic_to_shared (&sPre[0][FUSE ? threadIdx.x : 0]);
        auto fetch = [&] (unsigned kinds, int j, unsigned a, float4 &x, unsigned pad)
            {
            const unsigned kind = (kinds >> (4*j)) & 15u;
            if (kind == NUC_LOAD)
                x = partials4[tOff + a];
            else if (FUSE && kind == NUC_PRE)
                x = lds128 (sPreT + ((pad >> (4*j)) & 15u) * (NT * 16));
            else if (kind & NUC_TIP)
                {
                const unsigned t = (kinds >> (13 + 6*j)) & 63u;
                unsigned mask;
                asm volatile ("ld.shared.u8 %0, [%1];" : "=r"(mask) : "r"(sMaskP + t * PPB));
                x = lds128 (sTabK + (t * 16 + mask) * (K * 16));
                }
            };
        auto operand = [&] (unsigned kinds, int j, unsigned sp, const float4 &x) -> float4
            {
            if ((kinds >> (4*j)) & NUC_TIP)
                return x;
            return matvec4s (sPk + sp, x);
            };
        auto node = [&] (int oo, const float4 (&xi)[3], float4 (&xo)[3])
            {
            const uint4 oa = reinterpret_cast<const uint4 *>(nops + oo)[0];     // a1 a2 a3 kinds
            const uint4 ob = reinterpret_cast<const uint4 *>(nops + oo)[1];     // destOff sp1 sp2 sp3
            const unsigned kinds = oa.w;
            unsigned nk = 0;
            if (oo + 1 < nOp)
                {
                const uint4 na = reinterpret_cast<const uint4 *>(nops + oo + 1)[0];
                nk = na.w;
                const unsigned npad = FUSE ? (unsigned) nops[oo + 1].pad : 0u;
                fetch (nk, 0, na.x, xo[0], npad);
                fetch (nk, 1, na.y, xo[1], npad);
                if (nk & 0xf00u)
                    fetch (nk, 2, na.z, xo[2], npad);
                }
            float4 res = operand (kinds, 0, ob.y, xi[0]);
            float4 v   = operand (kinds, 1, ob.z, xi[1]);
            res.x *= v.x; res.y *= v.y; res.z *= v.z; res.w *= v.w;
            if (kinds & 0xf00u)                   // unrooted interior root: third neighbour
                {
                v = operand (kinds, 2, ob.w, xi[2]);
                res.x *= v.x; res.y *= v.y; res.z *= v.z; res.w *= v.w;
                }
            float m = 0.0f;                       // 0 marks "node not rescaled"
            if (kinds & NUC_RESCALE)
                {
                // lanes beyond K (K not a power of two) hold a copy of category K-1: harmless in a max
                m = fmaxf (fmaxf (fmaxf (res.x, res.y), fmaxf (res.z, res.w)), 0.0f);
                #pragma unroll
                for (int off = 1; off < L; off <<= 1)
                    m = fmaxf (m, __shfl_xor_sync (0xffffffffu, m, off));
                scale4 (res, m);
                }
            if (lk == 0)
                sNewT[oo * PPB] = m;
            if (active)
                partials4[tOff + ob.x] = res;
            if (nk & 0x888u)                      // the next node consumes this result (uniform test)
                {
                if (nk & NUC_FWD)         xo[0] = res;
                if (nk & (NUC_FWD << 4))  xo[1] = res;
                if (nk & (NUC_FWD << 8))  xo[2] = res;
                }
            cur = res;
            if (ci == 0) MB200_STAMP (8 + oo);
            };
        if (nOp > 0)
            {
            const uint4 na = reinterpret_cast<const uint4 *>(nops)[0];
            const unsigned npad = FUSE ? (unsigned) nops[0].pad : 0u;
            fetch (na.w, 0, na.x, xa[0], npad);
            fetch (na.w, 1, na.y, xa[1], npad);
            if (na.w & 0xf00u)
                fetch (na.w, 2, na.z, xa[2], npad);
            if (na.w & NUC_FWD)        xa[0] = cur;       // result of the previous chunk's last node
            if (na.w & (NUC_FWD << 4)) xa[1] = cur;
            if (na.w & (NUC_FWD << 8)) xa[2] = cur;
            }
        for (int oo = 0; oo < nOp; oo += 2)
            {
            node (oo, xa, xb);
            if (oo + 1 < nOp)
                node (oo + 1, xb, xa);
            }

        // ---- 4. batched scaler pass: logs with full ILP, node scalers out, old scalers in.  The L lanes
        //      of a pattern share its nodes, so only warp-level synchronisation is needed ----
        __syncwarp ();
        for (int oo = lk; oo < nOp; oo += L)
            {
            const int sw = nops[oo].sw, sr = nops[oo].sr;
            float sc = 0.0f, old = 0.0f;
            if (c < C)
                {
                if (sw >= 0)
                    {
                    // (float) log (double): CondLikeScaler_NUC4 / _SSE (src/likelihood.c:5183, 5328);
                    // the correctly rounded value, which the AVX variant's logf returns too in all
                    // but rare last-bit cases
                    sc = log_of_max (sNewT[oo * PPB]);
                    ctx.scalers[(size_t)sw * C + c] = sc;
                    if ((sEv.flags & MB200_GUARD_FLAG) && sc < MB200_GUARD_LN)
                        abortAcc = 1;                   // sparse rescaling ran this subtree too close to the float range
                    }
                if (sr >= 0)
                    old = ctx.scalers[(size_t)sr * C + c];
                }
            sNewT[oo * PPB] = sc;
            sOld[oo][pl] = old;
            }
        __syncwarp ();
        // site scaler: the reference's sequence  ... - old(o) + new(o) ...  (RemoveNodeScalers then
        // CondLikeScaler per node, src/likelihood.c:7938-7965), replayed per pattern
        if (lk == 0)
            for (int oo = 0; oo < nOp; oo++)
                {
                lnScaler -= sOld[oo][pl];
                if (nops[oo].sw >= 0)
                    lnScaler += sNewT[oo * PPB];
                }
        // ---- 5. after the last chunk: site scalers out, root integration, this tile's lnL terms ----
        if (ci < nChunk - 1)
            continue;
        {
        if (sEv.siteDst >= 0 && active && lk == 0)
            ctx.scalers[(size_t)sEv.siteDst * C + c] = lnScaler;

        if (sEv.root < 0)
            continue;

        // ---- root integration (Likelihood_NUC4_FMA, src/likelihood.c:6468-6625) ----
        if (!sEv.rootFwd)
            cur = partials4[tOff + sEv.rootOff];
        const double *freqs = sD + 2*K, *catW = sD + K;
        const float fA = (float) freqs[0], fC = (float) freqs[1], fG = (float) freqs[2], fT = (float) freqs[3];
        // the reference accumulates one fused chain over k = 0..K-1 and the four states; the chain
        // hops from lane to lane so that the rounding sequence is the same
        float likeF = 0.0f;
        if (sEv.equalWeights)
            {
            #pragma unroll
            for (int k = 0; k < K; k++)
                {
                float mine = fmaf (cur.x, fA, likeF);
                mine = fmaf (cur.y, fC, mine);
                mine = fmaf (cur.z, fG, mine);
                mine = fmaf (cur.w, fT, mine);
                likeF = __shfl_sync (0xffffffffu, mine, groupBase + k);
                }
            likeF *= (float) catW[0];
            }
        else
            {
            float s = cur.x * fA;
            s = fmaf (cur.y, fC, s);
            s = fmaf (cur.z, fG, s);
            s = fmaf (cur.w, fT, s);
            #pragma unroll
            for (int k = 0; k < K; k++)
                {
                const float mine = fmaf (s, (float) catW[kk], likeF);
                likeF = __shfl_sync (0xffffffffu, mine, groupBase + k);
                }
            }
        double likeI = 0.0;
        if (sEv.hasPInvar)
            {
            const unsigned int im = (unsigned int) ctx.invMask[cc];
            float li = (im & 1) ? fA : 0.0f;
            li = fmaf ((im & 2) ? 1.0f : 0.0f, fC, li);
            li = fmaf ((im & 4) ? 1.0f : 0.0f, fG, li);
            li = fmaf ((im & 8) ? 1.0f : 0.0f, fT, li);
            li *= (float) sEv.pInvar;
            likeI = (double) li;
            }
        int    abortFlag = 0;
        double term = 0.0;
        if (active && lk == 0)
            {
            term = site_term ((double) likeF, likeI, sEv.hasPInvar, sEv.flags & MB200_QUIRK_FLAG, lnScaler,
                              ctx.weights[(size_t)sEv.weightsRow * C + c], abortFlag);
            if ((sEv.flags & MB200_GUARD_FLAG) && likeF < MB200_GUARD_MIN)
                abortFlag = 1;
            }
        termAcc += term; abortAcc |= abortFlag;
        }
        }   // tiles of this CTA
        }   // chunks

    MB200_STAMP (4);
    if (sEv.root < 0)
        return;
    MB200_STAMP (5);
    finish_lnl<NT> (ctx, blockIdx.y, termAcc, abortAcc, out, seq);
    MB200_STAMP (6);
}

// ---- kernel entry points of the 4-state path ----
#ifndef MB200_FUSE_CTAS
#define MB200_FUSE_CTAS 2          // resident CTAs per SM the latency-path variants are compiled for
#endif
// job descriptors in global memory (device-resident batches, large jobs)
template <int K, int NT, bool FUSE>
__global__ void __launch_bounds__(NT, FUSE ? MB200_FUSE_CTAS : NUC_STREAM_THREADS / NT)
eval_nuc4_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
                  const DevChunk *__restrict__ chunks, const DevMat *__restrict__ cmats,
                  const DevOp *__restrict__ ops, DevResult *out, int seq, const __grid_constant__ JobIndex jx)
{
    nuc4_body<K, NT, FUSE> (ctx, evals, dvals, chunks, cmats, ops, out, seq, jx);
}

// job descriptors delivered in the kernel parameter block (host call path of small evaluations):
// no host->device copy on the way in
template <int K, int NT, int CAP>
__global__ void __launch_bounds__(NT, MB200_FUSE_CTAS)
eval_nuc4_pkernel (DevCtx ctx, BlobOffsets off, DevResult *out, int seq, const __grid_constant__ JobIndex jx,
                   const __grid_constant__ ParamBlob<CAP> blob)       // small uniform parameters first: they share the
                                                                       // constant-cache lines the kernel touches anyway
{
    const char *b = blob.bytes;
    nuc4_body<K, NT, true> (ctx, reinterpret_cast<const DevEval *>(b + off.eval), reinterpret_cast<const double *>(b + off.dbl),
                            reinterpret_cast<const DevChunk *>(b + off.chunk), reinterpret_cast<const DevMat *>(b + off.cmat),
                            reinterpret_cast<const DevOp *>(b + off.op), out, seq, jx);
}

// ---------------------------------------------------------------------------------------
// any S: fused evaluation on CUDA cores.  grid = (tiles of TP patterns, evaluations),
// block = NT.  Shared memory: one P matrix (S*S), one child tile (TP x (Sp+1)), the running
// product of the node (K*TP x S), per-pattern max and site scalers.  This is the correctness
// path for every state count; the 20- and 61-state tensor-core kernels take over where the
// update is a dense contraction.
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
eval_gen_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
                 const DevOp *__restrict__ ops, DevResult *out, int seq)
{
    extern __shared__ __align__(16) float smem[];
    const int S = ctx.S, Sp = ctx.Sp, K = ctx.K, C = ctx.C, TP = ctx.tilePatterns;
    const int KB = ctx.genKB;                     // rate categories per pass (all of them when their P matrices fit)
    const int ldc = Sp + 1, ldp = S + 1;          // odd leading dimensions: thread = (pattern, state i) reads row i of P conflict-free
    float *sPm   = smem;                          // [KB][S][ldp]
    float *sCh   = sPm + (size_t)KB*S*ldp;        // [KB][TP][ldc]
    float *sProd = sCh + (size_t)KB*TP*ldc;       // [K][TP][S]
    float *sMax  = sProd + (size_t)K*TP*S;        // [TP]
    float *sSite = sMax + TP;                     // [TP]
    int   *sFull = reinterpret_cast<int *>(sSite + TP);   // [TP] tip shortcut: pattern is missing
    const uint64_t fullMask = (S == 64) ? ~(uint64_t)0 : ((((uint64_t)1) << S) - 1);

    const DevEval *ev = evals + blockIdx.y;
    const double *catW = dvals + ev->dOff + K, *freqs = dvals + ev->dOff + 2*K;
    const int   c0 = blockIdx.x * TP;
    const int   np = min (TP, C - c0);            // patterns in this tile
    const size_t bufStride = (size_t)K * C * Sp;  // floats per partials buffer
    const int   lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarp = NT/32;

    for (int p = threadIdx.x; p < TP; p += NT)
        sSite[p] = (p < np && ev->siteSrc >= 0) ? ctx.scalers[(size_t)ev->siteSrc * C + c0 + p] : 0.0f;

    for (int o = 0; o < ev->nOp; o++)
        {
        const DevOp op = ops[ev->opOff + o];
        const int nChild = (op.c3 >= 0) ? 3 : 2;
        for (int ch = 0; ch < nChild; ch++)
            {
            const int child = (ch == 0) ? op.c1 : (ch == 1) ? op.c2 : op.c3;
            const int mat   = (ch == 0) ? op.m1 : (ch == 1) ? op.m2 : op.m3;
            const bool isTip = child < ctx.tipCount;
            const bool shortcut = isTip && (ev->flags & MB200_SHORTCUT_FLAG) && !ctx.tipPartAmbig[child];
            // one pass = one global round trip and two barriers: KB categories' P matrices and child rows at a time
            for (int k0 = 0; k0 < K; k0 += KB)
                {
                const int kb = min (KB, K - k0);
                __syncthreads ();
                const float *P = ctx.matrices + ((size_t)mat * K + k0) * S * S;
                for (int idx = threadIdx.x; idx < kb*S*S; idx += NT)
                    sPm[(idx / S) * ldp + idx % S] = P[idx];            // row (kk, i) of the pass at (kk*S + i) * ldp
                if (isTip)
                    {
                    for (int idx = threadIdx.x; idx < np*S; idx += NT)
                        {
                        const int p = idx / S, j = idx % S;
                        const uint64_t m = ctx.tip64[(size_t)child * C + c0 + p];
                        sCh[p*ldc + j] = ((m >> j) & 1) ? 1.0f : 0.0f;  // the same for every category
                        if (j == 0)
                            sFull[p] = (shortcut && m == fullMask) ? 1 : 0;
                        }
                    }
                else
                    {
                    for (int kk = 0; kk < kb; kk++)
                        {
                        const float *src = ctx.partials + (size_t)(child - ctx.tipCount) * bufStride
                                         + ((size_t)(k0 + kk) * C + c0) * Sp;
                        float *dstc = sCh + (size_t)kk * TP * ldc;
                        for (int idx = threadIdx.x; idx < np*Sp; idx += NT)
                            {
                            const int p = idx / Sp, j = idx % Sp;
                            if (j < S)
                                dstc[p*ldc + j] = src[idx];
                            }
                        }
                    }
                __syncthreads ();
                for (int idx = threadIdx.x; idx < kb*np*S; idx += NT)
                    {
                    const int kk = idx / (np*S), r = idx % (np*S);
                    const int p = r / S, i = r % S;
                    const float *prow = sPm + ((size_t)kk*S + i)*ldp;
                    const float *crow = sCh + (isTip ? 0 : (size_t)kk*TP*ldc) + p*ldc;
                    float acc = 0.0f;
                    for (int j = 0; j < S; j++)
                        acc = fmaf (prow[j], crow[j], acc);
                    if (isTip && sFull[p])
                        acc = 1.0f;                 // preLike shortcut (src/likelihood.c:257-258)
                    float *dst = sProd + ((size_t)(k0 + kk)*TP + p)*S + i;
                    *dst = (ch == 0) ? acc : (*dst) * acc;
                    }
                }
            }
        __syncthreads ();

        // per-pattern scaler: remove old, rescale, add new (one warp per pattern)
        for (int p = warp; p < np; p += nWarp)
            {
            float site = sSite[p];
            if (op.sr >= 0)
                site -= ctx.scalers[(size_t)op.sr * C + c0 + p];
            float m = 1.0f;
            if (op.sw >= 0)
                {
                m = 0.0f;
                for (int e = lane; e < K*S; e += 32)
                    m = fmaxf (m, sProd[((size_t)(e / S)*TP + p)*S + (e % S)]);
                #pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                    m = fmaxf (m, __shfl_xor_sync (0xffffffffu, m, off));
                const float sc = (float) log ((double) m);
                if (lane == 0)
                    ctx.scalers[(size_t)op.sw * C + c0 + p] = sc;
                site += sc;
                }
            if (lane == 0)
                {
                sMax[p]  = m;
                sSite[p] = site;
                }
            }
        __syncthreads ();
        {
        float *dstBase = ctx.partials + (size_t)(op.dest - ctx.tipCount) * bufStride;
        const bool scale = (op.sw >= 0);
        for (int k = 0; k < K; k++)
            {
            float *dst = dstBase + ((size_t)k * C + c0) * Sp;
            for (int idx = threadIdx.x; idx < np*Sp; idx += NT)
                {
                const int p = idx / Sp, j = idx % Sp;
                float v = 0.0f;
                if (j < S)
                    {
                    v = sProd[((size_t)k*TP + p)*S + j];
                    if (scale)
                        v /= sMax[p];
                    }
                dst[idx] = v;
                }
            }
        }
        }
    __syncthreads ();

    if (ev->siteDst >= 0)
        for (int p = threadIdx.x; p < np; p += NT)
            ctx.scalers[(size_t)ev->siteDst * C + c0 + p] = sSite[p];

    if (ev->root < 0)
        return;

    // ---- root integration (Likelihood_Gen, src/likelihood.c:5764-5916); accumulation in
    //      double, which is at least as accurate as the reference's float/double variants ----
    const float *rootBase = ctx.partials + (size_t)(ev->root - ctx.tipCount) * bufStride;
    double term = 0.0; int abortFlag = 0;
    for (int p = warp; p < np; p += nWarp)
        {
        double like = 0.0;
        for (int e = lane; e < K*S; e += 32)
            {
            const int k = e / S, s = e % S;
            const float v = rootBase[((size_t)k * C + c0 + p) * Sp + s];
            like += (double) v * freqs[s] * catW[k];
            }
        double likeI = 0.0;
        if (ev->hasPInvar)
            {
            const uint64_t im = ctx.invMask[c0 + p];
            for (int s = lane; s < S; s += 32)
                if ((im >> s) & 1)
                    likeI += freqs[s];
            likeI *= ev->pInvar;
            }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            {
            like  += __shfl_xor_sync (0xffffffffu, like, off);
            likeI += __shfl_xor_sync (0xffffffffu, likeI, off);
            }
        if (lane == 0)
            term += site_term (like, likeI, ev->hasPInvar, ev->flags & MB200_QUIRK_FLAG, sSite[p],
                               ctx.weights[(size_t)ev->weightsRow * C + c0 + p], abortFlag);
        }
    finish_lnl<NT> (ctx, blockIdx.y, term, abortFlag, out, seq);
}
