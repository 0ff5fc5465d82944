// mb200_kernels_eigen.cuh -- eigensystems of reversible rate matrices on the device
//
// What it replaces: the host side of UpDateCijk (src/likelihood.c:10476-10760): GetEigens (src/utils.c:11201, a general
// real-matrix solver: balance, Hessenberg reduction, QR iterations, inverse by LU) followed by CalcCijk (src/utils.c:9734).
// For a 61-state codon model that is 1-2 ms of host time per kappa / omega / pi move (3-5 ms with three omega categories),
// an order of magnitude more than the likelihood evaluation the move is followed by on this engine.
//
// Every rate matrix MrBayes builds for the divisions the seam accepts is time reversible: pi_i q_ij = pi_j q_ji.  With
// D = diag(sqrt(pi)),  A = D Q D^-1 is symmetric; A = U L U^T (U orthogonal) gives
//       Q = V L V^-1,   V = D^-1 U,   V^-1 = U^T D,
// and P(t) = V e^{Lt} V^-1 does not depend on which eigenbasis was found, so the transition probabilities agree with the
// reference's to the rounding of the double-precision sums (the float-rounded P(t) differ in the last place at most).
//
// Solver: cyclic Jacobi with a round-robin ("chess tournament") ordering: a round holds N/2 disjoint index pairs, all
// rotated at once.  Two kernels:
//   eigen_rotations_kernel  one CTA per matrix, the matrix in shared memory.  A thread owns the 2 x 2 block (rows of
//                           pair P) x (columns of pair Q) and applies the row rotation of P and the column rotation of Q
//                           to it in registers: A <- J^T A J in ONE pass, two barriers per round.  It does not carry the
//                           eigenvectors along (that would double the shared-memory traffic that bounds a round); it logs
//                           every rotation (c, s) instead.
//   eigen_vectors_kernel    U = U0 J_1 J_2 ... : every ROW of U is independent of the others, so one warp per row replays
//                           the log with warp-level synchronisation only (~30 us for 9 sweeps at N = 62).
// Warm start: a proposal changes a rate matrix a little, so the eigenvectors U0 of the chain's current matrix nearly
// diagonalise the proposed one: A0 = U0^T A U0 starts the sweeps with small off-diagonal entries and Jacobi's quadratic
// convergence needs 3-5 sweeps instead of 8-9.  U0 is orthogonal to rounding (a product of plane rotations); the engine
// restarts from the identity after MB200_EIG_WARM_CHAIN warm starts in a row so that rounding cannot accumulate.
#pragma once
#include <cuda_runtime.h>

#define MB200_EIG_THREADS   1024
#define MB200_EIG_NMAX      64
#define MB200_EIG_LD        MB200_EIG_NMAX        // even: the round's pairs (r + k, r - k) and the diagonal fall on distinct banks
#define MB200_EIG_SWEEPS    24
#define MB200_EIG_WARM_CHAIN 32
#define MB200_EIG_LOG_DOUBLES ((size_t) MB200_EIG_SWEEPS * (MB200_EIG_NMAX - 1) * (MB200_EIG_NMAX / 2) * 2)   // per matrix

// A | U0 | T (the two products of the warm start)
static inline size_t eigen_smem_bytes () { return (size_t)3 * MB200_EIG_NMAX * MB200_EIG_LD * sizeof(double); }

// pair k of round r among N (even) indices: index N-1 stays, the others walk round a circle
__device__ __forceinline__ void eig_pair (int N, int r, int k, int &p, int &q)
{
    const int M = N - 1;
    if (k == 0) { p = M; q = r; }
    else        { p = (r + k) % M; q = (r - k + M) % M; }
}

// grid = eigen parts; Q: parts x S x S (row major), pi: S; U0: parts x N x N eigenvectors to start from, or nullptr.
// out: rotation log (c, s per round and pair), number of rounds, eigenvalues into the slot's c_ijk block
__global__ void __launch_bounds__(MB200_EIG_THREADS)
eigen_rotations_kernel (const double *__restrict__ Q, const double *__restrict__ pi, int S, const double *__restrict__ U0,
                        double2 *__restrict__ rotLog, int *__restrict__ nRounds, double *__restrict__ block, int *status)
{
    extern __shared__ __align__(16) double eigS[];
    __shared__ double sD[MB200_EIG_NMAX], sC[MB200_EIG_NMAX / 2], sSn[MB200_EIG_NMAX / 2];
    __shared__ double sScale;
    __shared__ int    sFlag;
    double (*A)[MB200_EIG_LD] = reinterpret_cast<double (*)[MB200_EIG_LD]>(eigS);
    const int tid = threadIdx.x, part = blockIdx.x;
    const int N = (S + 1) & ~1, H = N >> 1;
    const double *q = Q + (size_t)part * S * S;
    double2 *rlog = rotLog + (size_t)part * (MB200_EIG_LOG_DOUBLES / 2);

    if (tid < N)
        sD[tid] = (tid < S) ? sqrt (pi[tid]) : 1.0;
    __syncthreads ();
    for (int e = tid; e < N * N; e += MB200_EIG_THREADS)
        {
        const int i = e / N, j = e % N;
        double a = 0.0;
        if (i < S && j < S)
            a = 0.5 * (sD[i] * q[i*S + j] / sD[j] + sD[j] * q[j*S + i] / sD[i]);   // symmetric up to rounding; take the mean
        A[i][j] = a;
        }
    if (U0 != nullptr)
        {
        // A <- U0^T A U0, 2 x 2 outputs per thread
        double (*Us)[MB200_EIG_LD] = reinterpret_cast<double (*)[MB200_EIG_LD]>(eigS + (size_t)MB200_EIG_NMAX * MB200_EIG_LD);
        double (*T)[MB200_EIG_LD]  = reinterpret_cast<double (*)[MB200_EIG_LD]>(eigS + (size_t)2 * MB200_EIG_NMAX * MB200_EIG_LD);
        const double *u0 = U0 + (size_t)part * N * N;
        for (int e = tid; e < N * N; e += MB200_EIG_THREADS)
            Us[e / N][e % N] = u0[e];
        __syncthreads ();
        const int i2 = (tid / H) * 2, j2 = (tid % H) * 2;
        if (tid < H * H)
            {
            double t00 = 0.0, t01 = 0.0, t10 = 0.0, t11 = 0.0;       // T = A U0
            for (int k = 0; k < N; k++)
                {
                const double a0 = A[i2][k], a1 = A[i2 + 1][k], b0 = Us[k][j2], b1 = Us[k][j2 + 1];
                t00 = fma (a0, b0, t00); t01 = fma (a0, b1, t01); t10 = fma (a1, b0, t10); t11 = fma (a1, b1, t11);
                }
            T[i2][j2] = t00; T[i2][j2 + 1] = t01; T[i2 + 1][j2] = t10; T[i2 + 1][j2 + 1] = t11;
            }
        __syncthreads ();
        if (tid < H * H && i2 <= j2)
            {
            double t00 = 0.0, t01 = 0.0, t10 = 0.0, t11 = 0.0;       // A = U0^T T, upper blocks, mirrored
            for (int k = 0; k < N; k++)
                {
                const double a0 = Us[k][i2], a1 = Us[k][i2 + 1], b0 = T[k][j2], b1 = T[k][j2 + 1];
                t00 = fma (a0, b0, t00); t01 = fma (a0, b1, t01); t10 = fma (a1, b0, t10); t11 = fma (a1, b1, t11);
                }
            if (i2 == j2) { t01 = 0.5 * (t01 + t10); t10 = t01; }
            A[i2][j2] = t00; A[i2][j2 + 1] = t01; A[i2 + 1][j2] = t10; A[i2 + 1][j2 + 1] = t11;
            A[j2][i2] = t00; A[j2 + 1][i2] = t01; A[j2][i2 + 1] = t10; A[j2 + 1][i2 + 1] = t11;
            }
        }
    __syncthreads ();
    if (tid == 0)
        {
        double s = 0.0;
        for (int i = 0; i < S; i++)
            s = fmax (s, fabs (A[i][i]));
        sScale = s;
        }
    __syncthreads ();
    const double tolCount = 1e-14 * sScale, tolSkip = 1e-19 * sScale;

    // a thread's block does not change from round to round: (row pair, column pair); the indices of pair k in round r
    // are r + k and r - k on the circle of M = N - 1 indices (pair 0: the fixed index M and r), no table needed
    const int  M = N - 1;
    const int  bP = tid / H, bQ = tid % H;
    const bool hasA = tid < H * H;

    int  rounds = 0;
    bool converged = false;
    for (int sweep = 0; sweep < MB200_EIG_SWEEPS; sweep++)
        {
        if (tid == 0) sFlag = 0;
        __syncthreads ();
        for (int r = 0; r < M; r++, rounds++)
            {
            if (tid < H)
                {
                int p = r + tid, qq = r - tid;
                if (p >= M) p -= M;
                if (qq < 0) qq += M;
                if (tid == 0) p = M;
                const double g = A[p][qq];
                double c = 1.0, s = 0.0;
                if (fabs (g) > tolSkip)
                    {
                    // tan of the rotation angle, smaller root:  t = 2 g sgn(h) / (|h| + sqrt (h^2 + 4 g^2)),  h = a_qq - a_pp.
                    // t may be a few ulps off (the rotation then leaves a_pq ~ 1e-16 |a_pq| behind, which the block update
                    // computes rather than assumes); c and s must satisfy c^2 + s^2 = 1: s = t c with c = rsqrt (1 + t^2)
                    const double h = A[qq][qq] - A[p][p];
                    const double w = fma (h, h, 4.0 * g * g);
                    const double den = fma (w, rsqrt (w), fabs (h));
                    const double t = copysign (2.0 * g, (h < 0.0) ? -g : g) * __drcp_rn (den);
                    c = rsqrt (fma (t, t, 1.0));
                    s = t * c;
                    if (fabs (g) > tolCount) sFlag = 1;
                    }
                sC[tid] = c; sSn[tid] = s;
                rlog[(size_t)rounds * H + tid] = make_double2 (c, s);
                }
            __syncthreads ();
            // A <- J^T A J : one 2 x 2 block per thread, rows of pair bP, columns of pair bQ
            if (hasA)
                {
                int r0 = r + bP, r1 = r - bP, c0 = r + bQ, c1 = r - bQ;
                if (r0 >= M) r0 -= M;
                if (r1 < 0)  r1 += M;
                if (c0 >= M) c0 -= M;
                if (c1 < 0)  c1 += M;
                if (bP == 0) r0 = M;
                if (bQ == 0) c0 = M;
                const double cr = sC[bP], sr = sSn[bP], cc = sC[bQ], sc = sSn[bQ];
                const double x00 = A[r0][c0], x01 = A[r0][c1], x10 = A[r1][c0], x11 = A[r1][c1];
                const double y00 = cr * x00 - sr * x10, y01 = cr * x01 - sr * x11;       // rows:   r0' = c r0 - s r1
                const double y10 = sr * x00 + cr * x10, y11 = sr * x01 + cr * x11;       //         r1' = s r0 + c r1
                const double z00 = cc * y00 - sc * y01, z01 = sc * y00 + cc * y01;       // columns likewise
                const double z10 = cc * y10 - sc * y11, z11 = sc * y10 + cc * y11;
                A[r0][c0] = z00; A[r0][c1] = z01; A[r1][c0] = z10; A[r1][c1] = z11;
                }
            __syncthreads ();
            }
        const int any = sFlag;
        __syncthreads ();
        if (!any) { converged = true; break; }
        }
    if (tid == 0)
        {
        nRounds[part] = rounds;
        if (!converged && status != nullptr)
            { *status = 1; __threadfence_system (); }
        }
    // the padded index (odd S) never rotates: eigenvalues 0 .. S-1 are the matrix's
    double *lam = block + (size_t)part * (2*(size_t)S + (size_t)S*S*S);
    if (tid < S)
        {
        lam[tid] = A[tid][tid];
        lam[S + tid] = 0.0;
        }
}

// grid = (ceil (N / 8), eigen parts), block = 256: one warp per row of U.  U = U0 (or I) times every logged rotation;
// the log passes through shared memory in chunks of 32 rounds (every warp of the CTA replays the same rotations), the
// next chunk's loads are in flight while the current one is applied.  Writes U (for the next warm start) and the
// factors V = D^-1 U, V^-1 = U^T D
#define MB200_EIG_CHUNK 32
__global__ void __launch_bounds__(256)
eigen_vectors_kernel (const double *__restrict__ pi, int S, const double *__restrict__ U0, const double2 *__restrict__ rotLog,
                      const int *__restrict__ nRounds, double *__restrict__ Uout, double *__restrict__ factor)
{
    __shared__ double  sRow[8][MB200_EIG_NMAX];
    __shared__ double2 sLog[2][MB200_EIG_CHUNK * (MB200_EIG_NMAX / 2)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, part = blockIdx.y;
    const int N = (S + 1) & ~1, H = N >> 1, M = N - 1;
    const int i = blockIdx.x * 8 + warp;
    const bool live = i < N;
    const double2 *rlog = rotLog + (size_t)part * (MB200_EIG_LOG_DOUBLES / 2);
    double *row = sRow[warp];
    if (live)
        for (int k = lane; k < N; k += 32)
            row[k] = (U0 != nullptr) ? U0[((size_t)part * N + i) * N + k] : ((k == i) ? 1.0 : 0.0);
    const int R = nRounds[part];
    const int per = MB200_EIG_CHUNK * H;                         // log entries per chunk
    constexpr int LPT = MB200_EIG_CHUNK * (MB200_EIG_NMAX / 2) / 256;   // entries per thread and chunk (4)
    double2 nx[LPT];
    auto fetch = [&] (int chunk)
        {
        #pragma unroll
        for (int j = 0; j < LPT; j++)
            {
            const int eIdx = j * 256 + threadIdx.x;
            const size_t g = (size_t)chunk * per + eIdx;
            nx[j] = (eIdx < per && g < (size_t)R * H) ? rlog[g] : make_double2 (1.0, 0.0);
            }
        };
    auto stash = [&] (int buf)
        {
        #pragma unroll
        for (int j = 0; j < LPT; j++)
            {
            const int eIdx = j * 256 + threadIdx.x;
            if (eIdx < per) sLog[buf][eIdx] = nx[j];
            }
        };
    const int nChunk = (R + MB200_EIG_CHUNK - 1) / MB200_EIG_CHUNK;
    if (nChunk > 0) { fetch (0); stash (0); }
    __syncthreads ();
    int rin = 0;
    for (int ck = 0; ck < nChunk; ck++)
        {
        const int buf = ck & 1;
        if (ck + 1 < nChunk) fetch (ck + 1);                     // in flight while this chunk is applied
        const int r1 = min (MB200_EIG_CHUNK, R - ck * MB200_EIG_CHUNK);
        if (live)
            for (int r = 0; r < r1; r++)
                {
                if (lane < H)
                    {
                    int p = rin + lane, q = rin - lane;
                    if (p >= M) p -= M;
                    if (q < 0)  q += M;
                    if (lane == 0) p = M;
                    const double2 cs = sLog[buf][r * H + lane];
                    const double a0 = row[p], a1 = row[q];
                    row[p] = cs.x * a0 - cs.y * a1;
                    row[q] = cs.y * a0 + cs.x * a1;
                    }
                __syncwarp ();
                if (++rin == M) rin = 0;
                }
        else
            { rin = (rin + r1) % M; }
        if (ck + 1 < nChunk) stash (buf ^ 1);
        __syncthreads ();
        }
    if (!live)
        return;
    double *uo = Uout + ((size_t)part * N + i) * N;
    for (int k = lane; k < N; k += 32)
        uo[k] = row[k];
    if (i < S)
        {
        const double d = sqrt (pi[i]);
        double *V = factor + (size_t)part * 2 * S * S, *W = V + (size_t)S * S;
        for (int k = lane; k < S; k += 32)
            {
            V[(size_t)i * S + k] = row[k] / d;        // V[i][k]
            W[(size_t)k * S + i] = row[k] * d;        // V^-1[k][j = i]
            }
        }
}

// c[i][j][k] = V[i][k] V^-1[k][j] for every part of a slot (CalcCijk, src/utils.c:9734-9746); grid = (blocks, parts)
__global__ void cijk_parts_kernel (double *__restrict__ block, const double *__restrict__ factor, int S)
{
    const int part = blockIdx.y;
    const size_t n3 = (size_t)S*S*S;
    double *c = block + (size_t)part * (2*(size_t)S + n3) + 2*S;
    const double *V = factor + (size_t)part * 2 * S * S, *W = V + (size_t)S * S;
    for (size_t idx = blockIdx.x*(size_t)blockDim.x + threadIdx.x; idx < n3; idx += (size_t)gridDim.x*blockDim.x)
        {
        const int k = (int)(idx % S);
        const int j = (int)((idx / S) % S);
        const int i = (int)(idx / ((size_t)S*S));
        c[idx] = V[i*S + k] * W[k*S + j];
        }
}
