// Dev probe: how fast can a CTA pull ~1 KB of job description out of the kernel parameter block
// (constant bank) into shared memory, cold?  Variants of the staging loop; globaltimer ns.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
struct Blob { char bytes[30720]; };
__device__ __forceinline__ unsigned long long now () { unsigned long long t; asm volatile ("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k (const __grid_constant__ Blob blob, const char *gblob, int words, unsigned long long *out, int *sink)
{
    __shared__ __align__(16) int s[1024];
    const unsigned long long t0 = now ();
    const int base = blockIdx.y * words;            // this evaluation's block (in 4-byte words)
    const int *pb = reinterpret_cast<const int *>(blob.bytes) + base;
    if (MODE == 0)
        for (int w = threadIdx.x; w < words; w += 256) s[w] = pb[w];
    else if (MODE == 1)
        for (int w = threadIdx.x; w < words / 4; w += 256) reinterpret_cast<int4 *>(s)[w] = reinterpret_cast<const int4 *>(pb)[w];
    else if (MODE == 2)
        {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int w = warp; w < words / 4; w += 8)
            {
            const int4 v = reinterpret_cast<const int4 *>(pb)[w];      // uniform address per warp
            if (lane == 0) reinterpret_cast<int4 *>(s)[w] = v;
            }
        }
    else if (MODE == 3)
        {
        const int *gb = reinterpret_cast<const int *>(gblob) + base;
        for (int w = threadIdx.x; w < words / 4; w += 256) reinterpret_cast<int4 *>(s)[w] = reinterpret_cast<const int4 *>(gb)[w];
        }
    else if (MODE == 4)
        for (int w = threadIdx.x; w < words / 2; w += 256) reinterpret_cast<int2 *>(s)[w] = reinterpret_cast<const int2 *>(pb)[w];
    else if (MODE == 5)
        {   // one warp only, 4 B per lane, the others wait
        if (threadIdx.x < 32)
            for (int w = threadIdx.x; w < words; w += 32) s[w] = pb[w];
        }
    __syncthreads ();
    const unsigned long long t1 = now ();
    if (threadIdx.x == 0) { out[(blockIdx.y * gridDim.x + blockIdx.x) * 2] = t0; out[(blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = t1; }
    if (s[(threadIdx.x * 7) % words] == 0x7fffffff) *sink = 1;
}

int main ()
{
    Blob *h = (Blob *) malloc (sizeof(Blob));
    char *g; cudaMalloc (&g, sizeof(Blob));
    unsigned long long *out, hout[2 * 64]; cudaMalloc (&out, sizeof(hout));
    int *sink; cudaMalloc (&sink, 4);
    char *flush; cudaMalloc (&flush, 256 << 20);
    const int words = 256;     // 1 KB per evaluation
    dim3 grid (4, 8);
    for (int mode = 0; mode < 6; mode++)
        {
        double sum = 0, mx = 0; int n = 0;
        for (int rep = 0; rep < 12; rep++)
            {
            for (size_t i = 0; i < sizeof(Blob); i++) h->bytes[i] = (char)(rand ());
            cudaMemcpy (g, h, sizeof(Blob), cudaMemcpyHostToDevice);
            cudaMemset (flush, rep, 256 << 20);
            cudaDeviceSynchronize ();
            switch (mode)
                {
                case 0: k<0><<<grid, 256>>> (*h, g, words, out, sink); break;
                case 1: k<1><<<grid, 256>>> (*h, g, words, out, sink); break;
                case 2: k<2><<<grid, 256>>> (*h, g, words, out, sink); break;
                case 3: k<3><<<grid, 256>>> (*h, g, words, out, sink); break;
                case 4: k<4><<<grid, 256>>> (*h, g, words, out, sink); break;
                case 5: k<5><<<grid, 256>>> (*h, g, words, out, sink); break;
                }
            cudaDeviceSynchronize ();
            cudaMemcpy (hout, out, sizeof(hout), cudaMemcpyDeviceToHost);
            if (rep < 2) continue;
            for (int c = 0; c < 32; c++) { double d = (double)(hout[2*c+1] - hout[2*c]); sum += d; if (d > mx) mx = d; n++; }
            }
        printf ("mode %d: mean %.0f ns, max %.0f ns  (%s)\n", mode, sum / n, mx,
                mode == 0 ? "4 B per thread, param" : mode == 1 ? "16 B per thread, param" : mode == 2 ? "16 B warp-uniform, param" :
                mode == 3 ? "16 B per thread, global (cold L2)" : mode == 4 ? "8 B per thread, param" : "one warp, 4 B per lane, param");
        }
    printf ("last error: %s\n", cudaGetErrorString (cudaGetLastError ()));
    return 0;
}
