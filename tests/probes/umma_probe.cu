// umma_probe.cu -- development probe (not product): one CTA computes D[128][N] = A[128][Kp] * B[N][Kp]^T
// with tcgen05.mma kind::tf32 (optionally 3xTF32 split) from software-laid-out shared memory operands.
// Validates the descriptor / layout conventions of umma_common.cuh before the pruning kernel uses them.
#include "../../mrbayes_b200/csrc/umma_common.cuh"
#include <stdio.h>
using namespace umma;

template <int N, int KP>
__global__ void __launch_bounds__(128) probe_kernel (const float *A, const float *B, float *D, int split)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *sAhi = reinterpret_cast<float *>(smem);                 // 128 x KP
    float *sAlo = sAhi + 128 * KP;
    float *sBhi = sAlo + 128 * KP;                                 // N x KP
    float *sBlo = sBhi + N * KP;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmemBase;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0)
        tmem_alloc<64> (&tmemBase);
    if (tid == 0)
        { mbar_init (&bar, 1); mbar_fence_init (); }
    for (int idx = tid; idx < 128 * KP; idx += 128)
        {
        const int r = idx / KP, j = idx % KP;
        const float x = A[idx], hi = to_tf32 (x), lo = to_tf32 (x - hi);
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sAhi) + canon_off (r, j, 128)) = hi;
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sAlo) + canon_off (r, j, 128)) = lo;
        }
    for (int idx = tid; idx < N * KP; idx += 128)
        {
        const int r = idx / KP, j = idx % KP;
        const float x = B[idx], hi = to_tf32 (x), lo = to_tf32 (x - hi);
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sBhi) + canon_off (r, j, N)) = hi;
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sBlo) + canon_off (r, j, N)) = lo;
        }
    fence_async_smem ();
    fence_before_sync ();
    __syncthreads ();
    fence_after_sync ();
    const uint32_t tbase = tmemBase;

    if (tid == 0)
        {
        const uint32_t idesc = make_idesc_tf32 (128, N);
        const uint32_t lboA = (128 / 8) * 128, lboB = (N / 8) * 128, sbo = 128;
        bool acc = false;
        const int passes = split ? 3 : 1;
        for (int pass = 0; pass < passes; pass++)
            {
            const float *pa = (pass == 2) ? sAlo : sAhi;
            const float *pb = (pass == 1) ? sBlo : sBhi;
            for (int ks = 0; ks < KP / 8; ks++)
                {
                const uint64_t da = make_desc (smem_u32 (pa) + ks * 2 * lboA, lboA, sbo);
                const uint64_t db = make_desc (smem_u32 (pb) + ks * 2 * lboB, lboB, sbo);
                mma_tf32 (tbase, da, db, idesc, acc);
                acc = true;
                }
            }
        mma_commit (&bar);
        }
    mbar_wait (&bar, 0);
    fence_after_sync ();
    // warp w reads TMEM lanes 32w .. 32w+31; thread t holds row 32w + t
    for (int c0 = 0; c0 < N; c0 += 16)
        {
        float v[16];
        tmem_ld16 (tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; i++)
            D[(size_t)tid * N + c0 + i] = v[i];
        }
    fence_before_sync ();
    __syncthreads ();
    if (warp == 0)
        tmem_dealloc<64> (tbase);
}

extern "C" int umma_probe (const float *hA, const float *hB, float *hD, int N, int KP, int split)
{
    float *dA, *dB, *dD;
    cudaMalloc (&dA, 128 * KP * 4); cudaMalloc (&dB, N * KP * 4); cudaMalloc (&dD, 128 * N * 4);
    cudaMemcpy (dA, hA, 128 * KP * 4, cudaMemcpyHostToDevice);
    cudaMemcpy (dB, hB, N * KP * 4, cudaMemcpyHostToDevice);
    cudaMemset (dD, 0, 128 * N * 4);
    size_t smem = (size_t)(2 * 128 * KP + 2 * N * KP) * 4;
    if (N == 64 && KP == 64)
        {
        cudaFuncSetAttribute (probe_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        probe_kernel<64, 64><<<1, 128, smem>>> (dA, dB, dD, split);
        }
    else if (N == 32 && KP == 24)
        {
        cudaFuncSetAttribute (probe_kernel<32, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        probe_kernel<32, 24><<<1, 128, smem>>> (dA, dB, dD, split);
        }
    else
        return -2;
    cudaError_t e = cudaDeviceSynchronize ();
    if (e != cudaSuccess) { fprintf (stderr, "umma_probe: %s\n", cudaGetErrorString (e)); return -1; }
    cudaMemcpy (hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    cudaFree (dA); cudaFree (dB); cudaFree (dD);
    return 0;
}
