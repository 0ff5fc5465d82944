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

// ---------------------------------------------------------------------------------------
// K1: transition matrices.  grid = (matrix updates, K), block = 128.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tiprobs_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const DevMat *__restrict__ mats)
{
    __shared__ double sExp[MB200_DEV_MAX_STATES];
    const DevMat   mu = mats[blockIdx.x];
    const int      k  = blockIdx.y;
    const int      S  = ctx.S;
    const DevEval *ev = evals + mu.eval;
    const double   t  = mu.length * ev->rates[k];
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
            P[idx] = (float) ev->freqs[idx % S];
        return;
        }
    const double *lam = ctx.eigen + (size_t)mu.eigen * (2*(size_t)S + (size_t)S*S*S);
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
                                            double *lnLOut, int *statusOut)
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
        ctx.tilePartial[(size_t)evalIdx*ctx.numTiles + blockIdx.x] = s;
        ctx.tileAbort  [(size_t)evalIdx*ctx.numTiles + blockIdx.x] = a;
        __threadfence ();
        unsigned int t = atomicAdd (&ctx.ticket[evalIdx], 1u);
        sLast = (t == (unsigned int)ctx.numTiles - 1u);
        }
    __syncthreads ();
    if (!sLast)
        return;
    __threadfence ();
    // last CTA of this evaluation: fixed-order sum over the tiles
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
        lnLOut[evalIdx]    = ab ? -DBL_MAX : tot;
        statusOut[evalIdx] = ab ? 1 : 0;
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
__device__ __forceinline__ float dot4_fma (const float4 p, const float4 x)
{
    // same operation order as CondLikeDown_NUC4_FMA (src/likelihood.c:1149-1169)
    return fmaf (p.w, x.w, fmaf (p.z, x.z, fmaf (p.y, x.y, p.x * x.x)));
}

__device__ __forceinline__ float4 matvec4 (const float4 *rows, const float4 x)
{
    return make_float4 (dot4_fma (rows[0], x), dot4_fma (rows[1], x), dot4_fma (rows[2], x), dot4_fma (rows[3], x));
}

template <int K, int NT>
__global__ void __launch_bounds__(NT)
eval_nuc4_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const DevOp *__restrict__ ops,
                  double *lnLOut, int *statusOut)
{
    __shared__ float4 sP  [2][3][K][4];     // P rows for interior children (double-buffered by op parity)
    __shared__ float4 sLut[2][3][K][16];    // tip tables: sum of the P columns selected by a state mask

    const DevEval *ev = evals + blockIdx.y;
    const int   C      = ctx.C;
    const int   c      = blockIdx.x * NT + threadIdx.x;
    const bool  active = c < C;
    const int   cc     = active ? c : C - 1;
    const size_t bufStride = (size_t)K * C;                 // float4 per partials buffer
    float4 *partials4 = reinterpret_cast<float4 *>(ctx.partials);

    float  lnScaler = (ev->siteSrc >= 0) ? ctx.scalers[(size_t)ev->siteSrc * C + cc] : 0.0f;
    float4 cur[K];
    int    curBuf = -2;
    #pragma unroll
    for (int k = 0; k < K; k++) cur[k] = make_float4 (0.f, 0.f, 0.f, 0.f);

    const int nOp = ev->nOp;

    // stage P rows / tip tables of operation o's branches into shared-memory buffer o & 1
    auto stage = [&] (int o)
        {
        const DevOp op = ops[ev->opOff + o];
        const int   pb = o & 1;
        const int   nChild = (op.c3 >= 0) ? 3 : 2;
        for (int ch = 0; ch < nChild; ch++)
            {
            const int child = (ch == 0) ? op.c1 : (ch == 1) ? op.c2 : op.c3;
            const int mat   = (ch == 0) ? op.m1 : (ch == 1) ? op.m2 : op.m3;
            const float4 *P4 = reinterpret_cast<const float4 *>(ctx.matrices + (size_t)mat * K * 16);
            if (child < ctx.tipCount)
                {
                // scalar-kernel shortcut: a missing observation on a tip without partial
                // ambiguity is exactly 1.0 (preLike tables, src/likelihood.c:816-832)
                const bool shortcut = (ev->flags & MB200_SHORTCUT_FLAG) && !ctx.tipPartAmbig[child];
                for (int e = threadIdx.x; e < K*64; e += NT)
                    {
                    const int i = e & 3, mask = (e >> 2) & 15, k = e >> 6;
                    const float4 row = P4[k*4 + i];
                    float r = (mask & 1) ? row.x : 0.0f;          // ascending-j sum of the
                    if (mask & 2) r += row.y;                     // selected columns == dense
                    if (mask & 4) r += row.z;                     // 0/1 matvec, bit for bit
                    if (mask & 8) r += row.w;
                    if (shortcut && mask == 15) r = 1.0f;
                    reinterpret_cast<float *>(&sLut[pb][ch][k][mask])[i] = r;
                    }
                }
            else
                {
                for (int e = threadIdx.x; e < K*4; e += NT)
                    sP[pb][ch][e >> 2][e & 3] = P4[e];
                }
            }
        };

    if (nOp > 0)
        stage (0);
    for (int o = 0; o < nOp; o++)
        {
        const DevOp op = ops[ev->opOff + o];
        const int   pb = o & 1;
        const int   nChild = (op.c3 >= 0) ? 3 : 2;

        // staged data of op o visible; every thread is past op o-1, so buffer (o+1)&1 is free
        __syncthreads ();

        // ---- issue this node's child loads first (2K..3K independent 16-byte requests) ----
        float4 x[3][K];
        int    tmask[3];
        #pragma unroll
        for (int ch = 0; ch < 3; ch++)
            {
            tmask[ch] = -1;
            if (ch == 2 && nChild == 2)
                break;
            const int child = (ch == 0) ? op.c1 : (ch == 1) ? op.c2 : op.c3;
            if (child < ctx.tipCount)
                tmask[ch] = ctx.tip8[(size_t)child * C + cc] & 15;
            else if (child == curBuf)
                {
                #pragma unroll
                for (int k = 0; k < K; k++) x[ch][k] = cur[k];
                }
            else
                {
                const float4 *src = partials4 + (size_t)(child - ctx.tipCount) * bufStride + cc;
                #pragma unroll
                for (int k = 0; k < K; k++) x[ch][k] = src[(size_t)k * C];
                }
            }

        // ---- overlap: stage the NEXT node's matrices while those loads are in flight ----
        if (o + 1 < nOp)
            stage (o + 1);

        // ---- two (three at the unrooted interior root) matvecs and their product ----
        float4 res[K];
        #pragma unroll
        for (int ch = 0; ch < 3; ch++)
            {
            if (ch == 2 && nChild == 2)
                break;
            float4 v[K];
            if (tmask[ch] >= 0)
                {
                #pragma unroll
                for (int k = 0; k < K; k++)
                    v[k] = sLut[pb][ch][k][tmask[ch]];
                }
            else
                {
                #pragma unroll
                for (int k = 0; k < K; k++)
                    v[k] = matvec4 (sP[pb][ch][k], x[ch][k]);
                }
            if (ch == 0)
                {
                #pragma unroll
                for (int k = 0; k < K; k++) res[k] = v[k];
                }
            else
                {
                #pragma unroll
                for (int k = 0; k < K; k++)
                    {
                    res[k].x *= v[k].x; res[k].y *= v[k].y; res[k].z *= v[k].z; res[k].w *= v[k].w;
                    }
                }
            }

        // ---- scaler bookkeeping: remove the node's old scaler, rescale, add the new one ----
        if (op.sr >= 0)
            lnScaler -= ctx.scalers[(size_t)op.sr * C + cc];
        if (op.sw >= 0)
            {
            float m = 0.0f;
            #pragma unroll
            for (int k = 0; k < K; k++)
                m = fmaxf (fmaxf (fmaxf (m, res[k].x), fmaxf (res[k].y, res[k].z)), res[k].w);
            #pragma unroll
            for (int k = 0; k < K; k++)
                {
                res[k].x /= m; res[k].y /= m; res[k].z /= m; res[k].w /= m;
                }
            const float sc = (float) log ((double) m);
            if (active)
                ctx.scalers[(size_t)op.sw * C + c] = sc;
            lnScaler += sc;
            }
        if (active)
            {
            float4 *dst = partials4 + (size_t)(op.dest - ctx.tipCount) * bufStride + c;
            #pragma unroll
            for (int k = 0; k < K; k++) dst[(size_t)k * C] = res[k];
            }
        #pragma unroll
        for (int k = 0; k < K; k++) cur[k] = res[k];
        curBuf = op.dest;
        }

    if (ev->siteDst >= 0 && active)
        ctx.scalers[(size_t)ev->siteDst * C + c] = lnScaler;

    if (ev->root < 0)
        return;

    // ---- root integration (Likelihood_NUC4_FMA, src/likelihood.c:6468-6625) ----
    if (ev->root != curBuf)
        {
        const float4 *src = partials4 + (size_t)(ev->root - ctx.tipCount) * bufStride + cc;
        #pragma unroll
        for (int k = 0; k < K; k++) cur[k] = src[(size_t)k * C];
        }
    const float fA = (float) ev->freqs[0], fC = (float) ev->freqs[1], fG = (float) ev->freqs[2], fT = (float) ev->freqs[3];
    float likeF;
    if (ev->equalWeights)
        {
        likeF = 0.0f;
        #pragma unroll
        for (int k = 0; k < K; k++)
            {
            likeF = fmaf (cur[k].x, fA, likeF);
            likeF = fmaf (cur[k].y, fC, likeF);
            likeF = fmaf (cur[k].z, fG, likeF);
            likeF = fmaf (cur[k].w, fT, likeF);
            }
        likeF *= (float) ev->catW[0];
        }
    else
        {
        likeF = 0.0f;
        #pragma unroll
        for (int k = 0; k < K; k++)
            {
            float s = cur[k].x * fA;
            s = fmaf (cur[k].y, fC, s);
            s = fmaf (cur[k].z, fG, s);
            s = fmaf (cur[k].w, fT, s);
            likeF = fmaf (s, (float) ev->catW[k], likeF);
            }
        }
    double likeI = 0.0;
    if (ev->hasPInvar)
        {
        const unsigned int im = (unsigned int) ctx.invMask[cc];
        float li = (im & 1) ? fA : 0.0f;
        li = fmaf ((im & 2) ? 1.0f : 0.0f, fC, li);
        li = fmaf ((im & 4) ? 1.0f : 0.0f, fG, li);
        li = fmaf ((im & 8) ? 1.0f : 0.0f, fT, li);
        li *= (float) ev->pInvar;
        likeI = (double) li;
        }
    int    abortFlag = 0;
    double term = 0.0;
    if (active)
        term = site_term ((double) likeF, likeI, ev->hasPInvar, ev->flags & MB200_QUIRK_FLAG, lnScaler,
                          ctx.weights[(size_t)ev->weightsRow * C + c], abortFlag);
    finish_lnl<NT> (ctx, blockIdx.y, term, abortFlag, lnLOut, statusOut);
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
eval_gen_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const DevOp *__restrict__ ops,
                 double *lnLOut, int *statusOut)
{
    extern __shared__ __align__(16) float smem[];
    const int S = ctx.S, Sp = ctx.Sp, K = ctx.K, C = ctx.C, TP = ctx.tilePatterns;
    const int ldc = Sp + 1;
    float *sPm   = smem;                          // [S][S]
    float *sCh   = sPm + S*S;                     // [TP][ldc]
    float *sProd = sCh + TP*ldc;                  // [K][TP][S]
    float *sMax  = sProd + (size_t)K*TP*S;        // [TP]
    float *sSite = sMax + TP;                     // [TP]
    int   *sFull = reinterpret_cast<int *>(sSite + TP);   // [TP] tip shortcut: pattern is missing
    const uint64_t fullMask = (S == 64) ? ~(uint64_t)0 : ((((uint64_t)1) << S) - 1);

    const DevEval *ev = evals + blockIdx.y;
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
            for (int k = 0; k < K; k++)
                {
                __syncthreads ();
                const float *P = ctx.matrices + ((size_t)mat * K + k) * S * S;
                for (int idx = threadIdx.x; idx < S*S; idx += NT)
                    sPm[idx] = P[idx];
                const bool isTip = child < ctx.tipCount;
                const bool shortcut = isTip && (ev->flags & MB200_SHORTCUT_FLAG) && !ctx.tipPartAmbig[child];
                if (isTip)
                    {
                    for (int idx = threadIdx.x; idx < np*S; idx += NT)
                        {
                        const int p = idx / S, j = idx % S;
                        const uint64_t m = ctx.tip64[(size_t)child * C + c0 + p];
                        sCh[p*ldc + j] = ((m >> j) & 1) ? 1.0f : 0.0f;
                        if (j == 0)
                            sFull[p] = (shortcut && m == fullMask) ? 1 : 0;
                        }
                    }
                else
                    {
                    const float *src = ctx.partials + (size_t)(child - ctx.tipCount) * bufStride
                                     + ((size_t)k * C + c0) * Sp;
                    for (int idx = threadIdx.x; idx < np*Sp; idx += NT)
                        {
                        const int p = idx / Sp, j = idx % Sp;
                        if (j < S)
                            sCh[p*ldc + j] = src[idx];
                        }
                    }
                __syncthreads ();
                for (int idx = threadIdx.x; idx < np*S; idx += NT)
                    {
                    const int p = idx / S, i = idx % S;
                    const float *prow = sPm + i*S;
                    const float *crow = sCh + p*ldc;
                    float acc = 0.0f;
                    for (int j = 0; j < S; j++)
                        acc = fmaf (prow[j], crow[j], acc);
                    if (isTip && sFull[p])
                        acc = 1.0f;                 // preLike shortcut (src/likelihood.c:257-258)
                    float *dst = sProd + ((size_t)k*TP + p)*S + i;
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
            like += (double) v * ev->freqs[s] * ev->catW[k];
            }
        double likeI = 0.0;
        if (ev->hasPInvar)
            {
            const uint64_t im = ctx.invMask[c0 + p];
            for (int s = lane; s < S; s += 32)
                if ((im >> s) & 1)
                    likeI += ev->freqs[s];
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
    finish_lnl<NT> (ctx, blockIdx.y, term, abortFlag, lnLOut, statusOut);
}
