// mb200_kernels_eigen.cuh -- eigensystems of reversible rate matrices on the device
//
// What it replaces: the host side of UpDateCijk (src/likelihood.c:10476-10760): GetEigens (src/utils.c:11201, a general
// real-matrix solver: balance, Hessenberg reduction, QR iterations, inverse by LU) followed by CalcCijk (src/utils.c:9734).
// For a 61-state codon model that is 2.1 ms of host time per kappa / omega / pi move (5.3 ms with three omega categories),
// an order of magnitude more than the likelihood evaluation the move is followed by on this engine.
//
// Every rate matrix MrBayes builds for the divisions the seam accepts is time reversible: pi_i q_ij = pi_j q_ji.  With
// D = diag(sqrt(pi)),  A = D Q D^-1 is symmetric; A = U L U^T (U orthogonal) gives
//       Q = V L V^-1,   V = D^-1 U,   V^-1 = U^T D,
// and P(t) = V e^{Lt} V^-1 does not depend on which eigenbasis was found, so the transition probabilities agree with the
// reference's to the rounding of the double-precision sums (the float-rounded P(t) differ in the last place at most).
//
// Solver: cyclic Jacobi with a round-robin ("chess tournament") ordering: a round holds N/2 disjoint index pairs, all
// rotated at once.  One CTA per matrix, 1024 threads; the matrix and the accumulated rotations live in shared memory.
// A thread owns the 2 x 2 block (rows of pair P) x (columns of pair Q) and applies the row rotation of P and the
// column rotation of Q to it in registers: A <- J^T A J in ONE pass, two barriers per round.  Jacobi is backward stable
// and converges quadratically; 6-9 sweeps of N-1 rounds for N = 62.
#pragma once
#include <cuda_runtime.h>

#define MB200_EIG_THREADS   1024
#define MB200_EIG_NMAX      64
#define MB200_EIG_LD        (MB200_EIG_NMAX + 1)
#define MB200_EIG_SWEEPS    40

static inline size_t eigen_smem_bytes () { return (size_t)2 * MB200_EIG_NMAX * MB200_EIG_LD * sizeof(double); }

// pair k of round r among N (even) indices: index N-1 stays, the others walk round a circle
__device__ __forceinline__ void eig_pair (int N, int r, int k, int &p, int &q)
{
    const int M = N - 1;
    if (k == 0) { p = M; q = r; }
    else        { p = (r + k) % M; q = (r - k + M) % M; }
}

// grid = eigen parts; Q: parts x S x S (row major), pi: S;  out: factor = [V (S x S) | V^-1 (S x S)] per part,
// block = the slot's c_ijk block (lambda, imaginary parts = 0 written here; the c_ijk themselves by cijk_parts_kernel)
__global__ void __launch_bounds__(MB200_EIG_THREADS)
eigen_jacobi_kernel (const double *__restrict__ Q, const double *__restrict__ pi, int S,
                     double *__restrict__ factor, double *__restrict__ block, int *status)
{
    extern __shared__ __align__(16) double eigS[];
    __shared__ double sD[MB200_EIG_NMAX], sC[MB200_EIG_NMAX / 2], sSn[MB200_EIG_NMAX / 2];
    __shared__ int    sP[MB200_EIG_NMAX / 2], sQ[MB200_EIG_NMAX / 2];
    __shared__ double sScale;
    __shared__ int    sFlag;
    double (*A)[MB200_EIG_LD] = reinterpret_cast<double (*)[MB200_EIG_LD]>(eigS);
    double (*U)[MB200_EIG_LD] = reinterpret_cast<double (*)[MB200_EIG_LD]>(eigS + (size_t)MB200_EIG_NMAX * MB200_EIG_LD);
    const int tid = threadIdx.x, part = blockIdx.x;
    const int N = (S + 1) & ~1, H = N >> 1;
    const double *q = Q + (size_t)part * S * S;

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
        U[i][j] = (i == j) ? 1.0 : 0.0;
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

    int sweep = 0;
    for (; sweep < MB200_EIG_SWEEPS; sweep++)
        {
        if (tid == 0) sFlag = 0;
        __syncthreads ();
        for (int r = 0; r < N - 1; r++)
            {
            if (tid < H)
                {
                int p, qq;
                eig_pair (N, r, tid, p, qq);
                const double g = A[p][qq];
                double c = 1.0, s = 0.0;
                if (fabs (g) > tolSkip)
                    {
                    const double theta = 0.5 * (A[qq][qq] - A[p][p]) / g;
                    double t;
                    if (fabs (theta) > 1e100) t = 0.5 / theta;
                    else
                        {
                        t = 1.0 / (fabs (theta) + sqrt (theta * theta + 1.0));
                        if (theta < 0.0) t = -t;
                        }
                    c = 1.0 / sqrt (t * t + 1.0);
                    s = t * c;
                    if (fabs (g) > tolCount) sFlag = 1;
                    }
                sP[tid] = p; sQ[tid] = qq; sC[tid] = c; sSn[tid] = s;
                }
            __syncthreads ();
            // A <- J^T A J : one 2 x 2 block per thread, rows of pair P, columns of pair Qp
            for (int b = tid; b < H * H; b += MB200_EIG_THREADS)
                {
                const int P = b / H, Qp = b % H;
                const int r0 = sP[P], r1 = sQ[P], c0 = sP[Qp], c1 = sQ[Qp];
                const double cr = sC[P], sr = sSn[P], cc = sC[Qp], sc = sSn[Qp];
                const double x00 = A[r0][c0], x01 = A[r0][c1], x10 = A[r1][c0], x11 = A[r1][c1];
                const double y00 = cr * x00 - sr * x10, y01 = cr * x01 - sr * x11;       // rows:   r0' = c r0 - s r1
                const double y10 = sr * x00 + cr * x10, y11 = sr * x01 + cr * x11;       //         r1' = s r0 + c r1
                double z00 = cc * y00 - sc * y01, z01 = sc * y00 + cc * y01;             // columns likewise
                double z10 = cc * y10 - sc * y11, z11 = sc * y10 + cc * y11;
                if (P == Qp) { z01 = 0.0; z10 = 0.0; }                                   // the annihilated pair, exactly
                A[r0][c0] = z00; A[r0][c1] = z01; A[r1][c0] = z10; A[r1][c1] = z11;
                }
            // U <- U J
            for (int b = tid; b < N * H; b += MB200_EIG_THREADS)
                {
                const int i = b / H, Qp = b % H;
                const int c0 = sP[Qp], c1 = sQ[Qp];
                const double cc = sC[Qp], sc = sSn[Qp];
                const double u0 = U[i][c0], u1 = U[i][c1];
                U[i][c0] = cc * u0 - sc * u1;
                U[i][c1] = sc * u0 + cc * u1;
                }
            __syncthreads ();
            }
        const int any = sFlag;
        __syncthreads ();
        if (!any) break;
        }
    if (sweep >= MB200_EIG_SWEEPS && tid == 0 && status != nullptr)
        { *status = 1; __threadfence_system (); }

    // the padded index (odd S) never rotates: eigenpairs 0 .. S-1 are the matrix's
    double *V = factor + (size_t)part * 2 * S * S, *W = V + (size_t)S * S;
    for (int e = tid; e < S * S; e += MB200_EIG_THREADS)
        {
        const int a = e / S, b = e % S;
        V[e] = U[a][b] / sD[a];          // V[i = a][k = b]
        W[e] = U[b][a] * sD[b];          // V^-1[k = a][j = b]
        }
    double *lam = block + (size_t)part * (2*(size_t)S + (size_t)S*S*S);
    if (tid < S)
        {
        lam[tid] = A[tid][tid];
        lam[S + tid] = 0.0;
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
