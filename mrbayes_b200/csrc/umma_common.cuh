// umma_common.cuh -- the few Blackwell (sm_100a) primitives the tensor-core pruning kernel needs,
// hand-written as inline PTX: TMEM allocation, shared-memory matrix descriptors, tcgen05.mma
// (kind::tf32), tcgen05.commit / mbarrier, tcgen05.ld, 1-D bulk async copies (TMA engine).
//
// Shared-memory operand layout used throughout (K-major, no swizzle; "INTERLEAVE" canonical form):
// a tile X[R rows][Kp floats] is stored as 8-row x 16-byte core matrices,
//     byte_offset(r, j) = (j/4) * (R/8)*128  +  (r/8) * 128  +  (r%8) * 16  +  (j%4) * 4
// i.e. consecutive 8-row groups are 128 B apart (SBO) and consecutive 16-byte K chunks are
// (R/8)*128 B apart (LBO).  One tcgen05.mma of kind::tf32 consumes K = 8 floats = two chunks.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32 (const void *p)
{
    return (uint32_t) __cvta_generic_to_shared (p);
}

// byte offset of element (r, j) in the canonical K-major layout of a tile with R rows
__device__ __forceinline__ uint32_t canon_off (int r, int j, int R)
{
    return (uint32_t)((j >> 2) * (R >> 3) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (j & 3) * 4);
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) leading byte offset>>4, [32,46) stride byte offset>>4,
// [46,48) version = 1 (Blackwell), [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc (uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::tf32, FP32 accumulate, both operands K-major
// (cute::UMMA::InstrDescriptor): c_format[4,6)=1 (F32), a_format[7,10)=2 (TF32), b_format[10,13)=2,
// a_major bit15 = 0, b_major bit16 = 0, n_dim[17,23)=N>>3, m_dim[24,29)=M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32 (int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_tf32 (uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate)
{
    uint32_t acc = accumulate ? 1u : 0u;
    asm volatile (
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// make the MMAs issued so far arrive on an mbarrier when they complete
__device__ __forceinline__ void mma_commit (uint64_t *bar)
{
    asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                  :: "r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync () { asm volatile ("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync ()  { asm volatile ("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma, bulk copies)
__device__ __forceinline__ void fence_async_smem ()  { asm volatile ("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- TMEM ----
template <int COLS>
__device__ __forceinline__ void tmem_alloc (uint32_t *smem_result)       // one full warp
{
    asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                  :: "r"(smem_u32 (smem_result)), "n"(COLS) : "memory");
    asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc (uint32_t taddr)            // the same warp
{
    asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "n"(COLS) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld16 (uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                  : "r"(taddr) : "memory");
    asm volatile ("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float (r[i]);
}

// same, 32 columns, WITHOUT the wait: issue several, then tmem_ld_wait() once
__device__ __forceinline__ void tmem_ld32_nowait (uint32_t taddr, uint32_t *r)
{
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                  "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                    "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                  : "r"(taddr) : "memory");
}
// 16 columns, without the wait
__device__ __forceinline__ void tmem_ld16_nowait (uint32_t taddr, uint32_t *r)
{
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                  : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait () { asm volatile ("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init (uint64_t *bar, int count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init () { asm volatile ("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx (uint64_t *bar, uint32_t bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
    asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity)
{
    asm volatile (
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n"
        :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}

// ---- 1-D bulk async copies (TMA engine; no tensor map needed for contiguous tiles) ----
__device__ __forceinline__ void bulk_g2s (void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                  :: "r"(smem_u32 (smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g (void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile ("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                  :: "l"(gdst), "r"(smem_u32 (smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit ()       { asm volatile ("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all () { asm volatile ("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// round-to-nearest TF32 (10-bit mantissa) of an fp32 value, returned as fp32 bits
__device__ __forceinline__ float to_tf32 (float x)
{
    uint32_t r;
    asm ("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float (r);
}

} // namespace umma
