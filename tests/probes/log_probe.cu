// Dev probe: log_of_max(m) against (float) log ((double) m) over a dense sweep of positive floats.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "mb200_kernels.cuh"
__global__ void sweep (unsigned long long *mismatch, unsigned long long *count, float *worst)
{
    unsigned long long bad = 0, n = 0;
    // every 97th float bit pattern from 2^-100 to 2 (rescaler maxima live in (0, 1])
    const unsigned lo = 0x0d800000u, hi = 0x40000000u;
    for (unsigned long long b = lo + (blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x) * 97ull; b < hi;
         b += (unsigned long long) gridDim.x * blockDim.x * 97ull)
        {
        const float m = __uint_as_float ((unsigned) b);
        const float a = log_of_max (m), r = (float) log ((double) m);
        n++;
        if (__float_as_uint (a) != __float_as_uint (r)) { bad++; *worst = m; }
        }
    atomicAdd (mismatch, bad); atomicAdd (count, n);
}
int main ()
{
    unsigned long long *d, h[2] = {0, 0}; float *w, hw = 0;
    cudaMalloc (&d, 16); cudaMemset (d, 0, 16); cudaMalloc (&w, 4); cudaMemset (w, 0, 4);
    sweep<<<1184, 256>>> (d, d + 1, w);
    cudaMemcpy (h, d, 16, cudaMemcpyDeviceToHost); cudaMemcpy (&hw, w, 4, cudaMemcpyDeviceToHost);
    printf ("log_of_max: %llu arguments, %llu differ from (float) log ((double) m) (last: %.9g); %s\n", h[1], h[0], hw, cudaGetErrorString (cudaGetLastError ()));
    return 0;
}
