// mb200_kernels_tc.cuh -- tensor-core (tcgen05 / TMEM) pruning kernel for the state counts where a
// node update really is a dense contraction: 20-state amino-acid and 61-state codon models
// (CondLikeDown/Root_Gen*, CondLikeScaler_Gen*, Likelihood_Gen*; reference src/likelihood.c:204,
// 2152, 4939, 5764).
//
// Per node, per rate category, per child:   D[128 patterns][S] = CL_child[128][S] * P^T[S][S]
// is one 128 x NP x KP tcgen05.mma chain (kind::tf32, FP32 accumulate in TMEM), NP/KP = S padded to
// the MMA granularity (61 -> 64/64, 20 -> 32/24).  FP32 accuracy is recovered with the 3xTF32 split
//     x = hi + lo,  hi = rna_tf32(x),  lo = rna_tf32(x - hi):   A*B ~= Ahi*Bhi + (Ahi*Blo + Alo*Bhi)
// (plain TF32 would cost ~2e-4 per product; the split leaves ~7e-7, see tests/probes/umma_probe.cu).
// The large term and the two small correction terms go to SEPARATE TMEM accumulators so that the
// tensor core's truncating accumulation bias is paid on KP/8 steps only, and are added in FP32 (RN)
// in the epilogue.
//
// CTA = 128 threads = one tile of 128 site patterns; grid = (tiles, evaluations); like the other
// pruning kernels the CTA walks the evaluation's whole operation list for its tile, no inter-CTA sync.
//   operands  A (child CLs): loaded from HBM with 512-byte-coalesced LDG.128, split hi/lo in registers,
//             written to shared memory in the canonical K-major core-matrix layout (umma_common.cuh);
//             the child that is the previous node's result comes straight from registers; tips are
//             expanded from their state masks (exact in TF32, no lo term).
//             B (P(t), pre-split hi/lo in canonical layout by tiprobs_kernel): one bulk async copy
//             (TMA engine, mbarrier complete_tx) per child and category.
//   MMA       one elected thread issues 3 * KP/8 tcgen05.mma; tcgen05.commit -> mbarrier.
//   epilogue  thread t = pattern row t = TMEM lane t: tcgen05.ld, product over children in registers,
//             max / divide / log, coalesced-row float4 stores, site-scaler bookkeeping, root integration.
#pragma once
#include "mb200_device.cuh"
#include "umma_common.cuh"

template <int S> struct TcGeom;
template <> struct TcGeom<61> { static constexpr int NP = 64, KP = 64, KMAX = 1, SP = 64; };
template <> struct TcGeom<20> { static constexpr int NP = 32, KP = 24, KMAX = 4, SP = 20; };

// SLOTS: children whose operands are staged and whose MMAs are in flight at the same time.  Two slots
// (one barrier round, one MMA wait and one read-out per node: 28 % less time per CTA, measured on
// the 61-state tile) need twice the shared memory, i.e. one CTA per SM: the engine uses them for
// small grids (at most one CTA per SM anyway) and one slot -- two co-resident CTAs -- for large ones.
template <int S, int SLOTS> __host__ __device__ constexpr size_t tc_smem_bytes ()
{
    return (size_t) SLOTS * TcGeom<S>::KMAX * (2 * 128 * TcGeom<S>::KP + 2 * TcGeom<S>::NP * TcGeom<S>::KP) * sizeof(float);
}

// floats per pre-split matrix image: ONE canonical-layout operand of 2 NP rows x KP -- rows [0, NP) hold the hi
// parts, rows [NP, 2 NP) the lo parts, so that  A x [B_hi | B_lo]^T  is a single N = 2 NP MMA chain (the pipelined
// kernel) and B_hi / B_lo alone are the same image addressed with N = NP (row offset 0 / NP)
template <int S> __host__ __device__ constexpr int tc_split_floats () { return 2 * TcGeom<S>::NP * TcGeom<S>::KP; }

// P(t) [S][S] row-major (row = ancestral state i) -> hi/lo rows of B[n = i][k = j] in canonical layout
template <int S>
__device__ __forceinline__ void tc_write_split_entry (float *img, int i, int j, float p)
{
    constexpr int NP = TcGeom<S>::NP;
    const float hi = umma::to_tf32 (p), lo = umma::to_tf32 (p - hi);
    img[umma::canon_off (i, j, 2 * NP) / 4] = hi;
    img[umma::canon_off (NP + i, j, 2 * NP) / 4] = lo;
}

// split images of matrices already present in the matrix buffer (set_transition_matrix, or a
// tiprobs launch predating the split buffer): grid = (matrices, K)
template <int S>
__global__ void tc_split_kernel (const float *__restrict__ matrices, float *__restrict__ split, const DevMat *__restrict__ upd,
                                 int first, int K)
{
    constexpr int NP = TcGeom<S>::NP, KP = TcGeom<S>::KP;
    const int m = upd ? upd[blockIdx.x].matrix : first + blockIdx.x, k = blockIdx.y;
    const float *P = matrices + ((size_t)m * K + k) * S * S;
    float *img = split + ((size_t)m * K + k) * tc_split_floats<S> ();
    for (int idx = threadIdx.x; idx < NP * KP; idx += blockDim.x)
        {
        const int i = idx / KP, j = idx % KP;
        tc_write_split_entry<S> (img, i, j, (i < S && j < S) ? P[i * S + j] : 0.0f);
        }
}

template <int S, int SLOTS>
__global__ void __launch_bounds__(128, 1)
eval_tc_kernel (DevCtx ctx, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
                const DevOp *__restrict__ ops, const float *__restrict__ split, DevResult *out, int seq)
{
    using namespace umma;
    constexpr int NP = TcGeom<S>::NP, KP = TcGeom<S>::KP, KMAX = TcGeom<S>::KMAX;
    constexpr int TM = 128;                                   // patterns per tile = MMA M
    constexpr int ACC_COLS = 2 * NP * KMAX;                   // TMEM columns of one child's accumulators
    constexpr int TMEM_COLS = (SLOTS * ACC_COLS <= 64) ? 64 : (SLOTS * ACC_COLS <= 128) ? 128 : (SLOTS * ACC_COLS <= 256) ? 256 : 512;
    constexpr uint32_t LBO_A = (TM / 8) * 128, LBO_B = (2 * NP / 8) * 128, SBO = 128;
    constexpr int A_FLOATS = TM * KP;                         // one hi (or lo) image
    constexpr int B_FLOATS = 2 * NP * KP;                     // hi + lo image of one P(t)
    constexpr int NQ = (S + 3) / 4;                           // 16-byte chunks per stored row
    constexpr int SPC = TcGeom<S>::SP;                        // floats per global row (ctx.Sp)
    constexpr int A_SLOT = KMAX * 2 * A_FLOATS;               // floats of one child's A images
    constexpr int B_SLOT = KMAX * B_FLOATS;

    // dynamic shared memory: [slot][k][hi|lo] A images, then [slot][k] B images; the first A slot
    // doubles as the staging area of the node's result rows once the MMAs have completed
    extern __shared__ __align__(128) unsigned char tc_smem[];
    float *sA = reinterpret_cast<float *>(tc_smem);           // SLOTS x A_SLOT
    float *sB = sA + SLOTS * A_SLOT;                          // SLOTS x B_SLOT
    float4 *sStage = reinterpret_cast<float4 *>(tc_smem);     // [k][q][TM+1] float4
    __shared__ uint64_t barB, barM;
    __shared__ uint32_t tmemBase;
    __shared__ DevEval sEv;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sp = ctx.Sp, K = ctx.K, C = ctx.C;
    if (tid < (int)(sizeof(DevEval) / 4))
        reinterpret_cast<int *>(&sEv)[tid] = reinterpret_cast<const int *>(evals + blockIdx.y)[tid];
    if (warp == 0)
        tmem_alloc<TMEM_COLS> (&tmemBase);
    // warps whose leader lane issues MMAs (one slot: the 2K accumulators of a child; two slots: all four
    // leaders arrive on the barrier, with or without work)
    const int nIssuers = (SLOTS == 1) ? ((2 * K < 4) ? 2 * K : 4) : 4;
    if (tid == 0)
        { mbar_init (&barB, 1); mbar_init (&barM, nIssuers); mbar_fence_init (); }
    fence_before_sync ();
    __syncthreads ();
    fence_after_sync ();
    const uint32_t tBase = tmemBase;                           // [slot][k][main | corr] accumulators, NP columns each
    const uint32_t laneSel = (uint32_t)(warp * 32) << 16;      // this warp's TMEM lane quadrant
    uint32_t parB = 0, parM = 0;

    const double *catW = dvals + sEv.dOff + K, *freqs = dvals + sEv.dOff + 2*K;
    // a tile may hold fewer than the MMA's 128 rows (ctx.tilePatterns, a multiple of 8): more, smaller CTAs
    // co-resident per SM overlap each other's load / MMA / read-out phases; the unused rows are zero padding
    const int   c0 = blockIdx.x * ctx.tilePatterns;
    const int   np = min (ctx.tilePatterns, C - c0);
    const int   c  = c0 + tid;                                 // this thread's pattern (epilogue role)
    const bool  active = tid < np;
    const size_t bufStride = (size_t)K * C * Sp;
    const bool  shortcutFlag = (sEv.flags & MB200_SHORTCUT_FLAG) != 0;
    const uint64_t fullMask = (S == 64) ? ~(uint64_t)0 : ((((uint64_t)1) << S) - 1);
    constexpr uint32_t idesc = make_idesc_tf32 (TM, NP);

    float site = (active && sEv.siteSrc >= 0) ? ctx.scalers[(size_t)sEv.siteSrc * C + c] : 0.0f;

    // operand descriptors never change: images live at fixed shared-memory addresses.  Only the
    // start-address field (bits 0..13, 16-byte units) moves with the slot, the category and the K step.
    const uint64_t dA0 = make_desc (smem_u32 (sA), LBO_A, SBO);      // hi image of slot 0, category 0
    const uint64_t dB0 = make_desc (smem_u32 (sB), LBO_B, SBO);
    constexpr uint64_t A_LO = (A_FLOATS * 4) >> 4, A_K = (2 * A_FLOATS * 4) >> 4, A_KS = (2 * LBO_A) >> 4, A_SL = ((uint64_t) A_SLOT * 4) >> 4;
    constexpr uint64_t B_LO = ((NP / 8) * 128) >> 4, B_K = (B_FLOATS * 4) >> 4, B_KS = (2 * LBO_B) >> 4, B_SL = ((uint64_t) B_SLOT * 4) >> 4;

    int preloaded = -1;          // partials buffer whose hi/lo images already sit in A slot 0 (previous node's result)

    // ---- helpers ----
    // B: the K pre-split P(t) images of a branch, one bulk async copy (contiguous) into slot `sl`
    auto stageB = [&] (int sl, int mat)
        {
        bulk_g2s (sB + sl * B_SLOT, split + (size_t)mat * K * B_FLOATS, (uint32_t)(K * B_FLOATS * 4), &barB);
        };
    // A: child tiles of all K categories -> hi / lo canonical images in slot `sl`; returns whether the
    // scalar-kernel tip shortcut applies to this thread's pattern
    auto stageA = [&] (int sl, int child, bool isTip) -> bool
        {
        float *slotA = sA + sl * A_SLOT;
        if (isTip)
            {
            // thread t expands pattern t's state mask (identical for every category): 0/1 are exact
            // in TF32, the lo image is not used
            const uint64_t m = active ? ctx.tip64[(size_t)child * C + c] : 0;
            #pragma unroll
            for (int q = 0; q < KP / 4; q++)
                {
                float4 h;
                h.x = ((m >> (q*4 + 0)) & 1) ? 1.f : 0.f; h.y = ((m >> (q*4 + 1)) & 1) ? 1.f : 0.f;
                h.z = ((m >> (q*4 + 2)) & 1) ? 1.f : 0.f; h.w = ((m >> (q*4 + 3)) & 1) ? 1.f : 0.f;
                *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(slotA) + canon_off (tid, q*4, TM)) = h;
                }
            return shortcutFlag && active && m == fullMask && !ctx.tipPartAmbig[child];
            }
        // HBM/L2 -> registers -> shared: a warp covers 8 rows x 4 chunks (64 contiguous bytes per
        // row: full 32-byte sectors) and stores 8 x 16 B contiguous per quarter-warp (no conflicts).
        // All loads of a category are issued before the first one is used (memory-level
        // parallelism: one round trip per category instead of one per chunk).
        constexpr int QB = (KP / 4 + 3) / 4;           // chunk blocks of 4
        constexpr int NIT = (TM / 8) * QB / 4;         // items per warp and category
        for (int k = 0; k < K; k++)
            {
            const float *src = ctx.partials + (size_t)(child - ctx.tipCount) * bufStride + ((size_t)k * C + c0) * Sp;
            unsigned char *base = reinterpret_cast<unsigned char *>(slotA + (size_t)k * 2 * A_FLOATS);
            float4 x[NIT];
            #pragma unroll
            for (int n = 0; n < NIT; n++)
                {
                const int it = warp + 4 * n;
                const int rb = it / QB, qb = it % QB;
                const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                x[n] = make_float4 (0.f, 0.f, 0.f, 0.f);
                if (q < KP / 4 && r < np && q * 4 < SPC)
                    x[n] = __ldcg (reinterpret_cast<const float4 *>(src + (size_t)r * SPC + q * 4));
                }
            #pragma unroll
            for (int n = 0; n < NIT; n++)
                {
                const int it = warp + 4 * n;
                const int rb = it / QB, qb = it % QB;
                const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                if (q < KP / 4)
                    {
                    const float4 h = make_float4 (to_tf32 (x[n].x), to_tf32 (x[n].y), to_tf32 (x[n].z), to_tf32 (x[n].w));
                    const float4 l = make_float4 (to_tf32 (x[n].x - h.x), to_tf32 (x[n].y - h.y), to_tf32 (x[n].z - h.z), to_tf32 (x[n].w - h.w));
                    *reinterpret_cast<float4 *>(base + canon_off (r, q*4, TM)) = h;
                    *reinterpret_cast<float4 *>(base + A_FLOATS * 4 + canon_off (r, q*4, TM)) = l;
                    }
                }
            }
        return false;
        };
    // MMA for the children staged in slots [0, nSl): main[k] = Ahi*Bhi ; corr[k] = Ahi*Blo (+ Alo*Bhi).
    // Issue is the bottleneck of these small MMAs (~35 ns each from one thread), so the independent
    // accumulators are spread over the four warps' leader lanes; every leader arrives on barM
    auto issueMMA = [&] (int nSl, unsigned tipBits)
        {
        if (lane != 0 || warp >= nIssuers)
            return;
        bool any = false;
        for (int item = warp; item < nSl * 2 * K; item += 4)
            {
            const int sl = (SLOTS == 1) ? 0 : item / (2 * K), k = (SLOTS == 1) ? (item >> 1) : (item >> 1) % K, corr = item & 1;
            const bool isTip = (tipBits >> sl) & 1u;
            const uint64_t aHi = dA0 + (uint64_t) sl * A_SL + (isTip ? 0 : (uint64_t)k * A_K), aLo = aHi + A_LO;
            const uint64_t bHi = dB0 + (uint64_t) sl * B_SL + (uint64_t)k * B_K, bLo = bHi + B_LO;
            const uint32_t tAcc = tBase + sl * ACC_COLS + k * 2 * NP + corr * NP;
            if (!corr)
                {
                #pragma unroll
                for (int ks = 0; ks < KP / 8; ks++)
                    mma_tf32 (tAcc, aHi + ks * A_KS, bHi + ks * B_KS, idesc, ks > 0);
                }
            else
                {
                #pragma unroll
                for (int ks = 0; ks < KP / 8; ks++)
                    mma_tf32 (tAcc, aHi + ks * A_KS, bLo + ks * B_KS, idesc, ks > 0);
                if (!isTip)
                    {
                    #pragma unroll
                    for (int ks = 0; ks < KP / 8; ks++)
                        mma_tf32 (tAcc, aLo + ks * A_KS, bHi + ks * B_KS, idesc, true);
                    }
                }
            any = true;
            }
        if (any) mma_commit (&barM);
        else     mbar_arrive (&barM);
        };

    for (int o = 0; o < sEv.nOp; o++)
        {
        const DevOp op = ops[sEv.opOff + o];
        const int nChild = (op.c3 >= 0) ? 3 : 2;
        float res[KMAX][S];                                    // this pattern's node result, all categories

        // my row of D for the child in slot `sl` (all categories), times what the other children gave
        auto readAcc = [&] (int sl, bool tipFull, bool firstChild)
            {
            #pragma unroll
            for (int k = 0; k < KMAX; k++)
                {
                if (k >= K) break;
                #pragma unroll
                for (int cb = 0; cb < NP; cb += 32)
                    {
                    uint32_t vm[32], vc[32];
                    tmem_ld32_nowait (tBase + sl * ACC_COLS + k * 2 * NP + laneSel + cb, vm);
                    tmem_ld32_nowait (tBase + sl * ACC_COLS + k * 2 * NP + NP + laneSel + cb, vc);
                    tmem_ld_wait ();
                    #pragma unroll
                    for (int i = 0; i < 32; i++)
                        if (cb + i < S)
                            {
                            float v = __uint_as_float (vm[i]) + __uint_as_float (vc[i]);
                            if (tipFull) v = 1.0f;             // preLike shortcut (src/likelihood.c:257-258)
                            res[k][cb + i] = firstChild ? v : res[k][cb + i] * v;
                            }
                    }
                }
            };

        // the child that is the previous node's result goes first: its images are already in slot 0
        int first = 0;
        if (preloaded >= 0)
            first = (op.c1 == preloaded) ? 0 : (op.c2 == preloaded) ? 1 : (op.c3 == preloaded) ? 2 : 0;
        for (int cc = 0; cc < nChild; cc += SLOTS)
            {
            const int nSl = (nChild - cc < SLOTS) ? nChild - cc : SLOTS;
            bool     tipFull[SLOTS];
            unsigned tipBits = 0;
            if (tid == 0)
                mbar_expect_tx (&barB, (uint32_t)(nSl * K * B_FLOATS * 4));
            #pragma unroll
            for (int sl = 0; sl < SLOTS; sl++)
                {
                tipFull[sl] = false;
                if (sl >= nSl) continue;
                const int ci = cc + sl;
                const int ch = (ci == 0) ? first : (ci <= first) ? ci - 1 : ci;
                const int child = (ch == 0) ? op.c1 : (ch == 1) ? op.c2 : op.c3;
                const int mat   = (ch == 0) ? op.m1 : (ch == 1) ? op.m2 : op.m3;
                const bool isTip = child < ctx.tipCount;
                if (tid == 0)
                    stageB (sl, mat);
                if (isTip) tipBits |= 1u << sl;
                if (!(ci == 0 && child == preloaded))          // else: images already in place (slot 0)
                    tipFull[sl] = stageA (sl, child, isTip);
                }
            fence_async_smem ();                               // generic-proxy stores -> async proxy (MMA)
            mbar_wait (&barB, parB); parB ^= 1;                // B images landed
            fence_before_sync ();
            __syncthreads ();
            fence_after_sync ();
            issueMMA (nSl, tipBits);
            mbar_wait (&barM, parM); parM ^= 1;
            fence_after_sync ();
            #pragma unroll
            for (int sl = 0; sl < SLOTS; sl++)
                if (sl < nSl)
                    readAcc (sl, tipFull[sl], cc + sl == 0);
            fence_before_sync ();                              // TMEM reads ordered before the next MMA
            __syncthreads ();                                  // shared operands free for the next children
            fence_after_sync ();
            }

        // ---- epilogue part 2: scaler bookkeeping and rescale (one thread = one pattern) ----
        if (active)
            {
            if (op.sr >= 0)
                site -= ctx.scalers[(size_t)op.sr * C + c];
            if (op.sw >= 0)
                {
                float m = 0.0f;
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    if (k < K)
                        {
                        #pragma unroll
                        for (int i = 0; i < S; i++) m = fmaxf (m, res[k][i]);
                        }
                // one IEEE reciprocal, then multiplies: 1 ulp from the reference's divisions, far below
                // the 3xTF32 operand error of this path
                const float rcp = 1.0f / m;
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    if (k < K)
                        {
                        #pragma unroll
                        for (int i = 0; i < S; i++) res[k][i] *= rcp;
                        }
                const float sc = (float) log ((double) m);     // CondLikeScaler_Gen_SSE, src/likelihood.c:5055
                ctx.scalers[(size_t)op.sw * C + c] = sc;
                site += sc;
                }
            }
        // ---- store: rows go through shared memory (chunk-major, conflict-free) so that the global
        //      writes are fully coalesced 16-byte-per-lane runs of the contiguous tile ----
        #pragma unroll
        for (int k = 0; k < KMAX; k++)
            if (k < K)
                {
                #pragma unroll
                for (int q = 0; q < NQ; q++)
                    {
                    float4 v;
                    v.x = res[k][q*4];
                    v.y = (q*4 + 1 < S) ? res[k][(q*4 + 1 < S) ? q*4 + 1 : 0] : 0.f;
                    v.z = (q*4 + 2 < S) ? res[k][(q*4 + 2 < S) ? q*4 + 2 : 0] : 0.f;
                    v.w = (q*4 + 3 < S) ? res[k][(q*4 + 3 < S) ? q*4 + 3 : 0] : 0.f;
                    sStage[((size_t)k * NQ + q) * (TM + 1) + tid] = v;
                    }
                }
        __syncthreads ();
        {
        float *dstBase = ctx.partials + (size_t)(op.dest - ctx.tipCount) * bufStride;
        constexpr int nq = SPC / 4;                            // chunks per global row (pad chunks are zero)
        for (int k = 0; k < K; k++)
            {
            float4 *dst = reinterpret_cast<float4 *>(dstBase + ((size_t)k * C + c0) * SPC);
            #pragma unroll
            for (int n = 0; n < nq; n++)                       // 128 rows x nq chunks = nq rounds of 128 lanes
                {
                const int idx = n * 128 + tid;
                const int r = idx / nq, q = idx % nq;
                if (r < np)
                    dst[idx] = (q < NQ) ? sStage[((size_t)k * NQ + q) * (TM + 1) + r] : make_float4 (0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncthreads ();                                      // staging area is A slot 0 again
        // register forwarding: when the next node consumes this result, its hi/lo images are written
        // straight from registers (no store -> load round trip through L2 on dependent chains)
        preloaded = -1;
        if (o + 1 < sEv.nOp)
            {
            const DevOp nx = ops[sEv.opOff + o + 1];
            if (nx.c1 == op.dest || nx.c2 == op.dest || nx.c3 == op.dest)
                {
                preloaded = op.dest;
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    if (k < K)
                        {
                        unsigned char *base = reinterpret_cast<unsigned char *>(sA + (size_t)k * 2 * A_FLOATS);
                        #pragma unroll
                        for (int q = 0; q < KP / 4; q++)
                            {
                            float x[4], h[4], l[4];
                            #pragma unroll
                            for (int e = 0; e < 4; e++)
                                {
                                x[e] = (q*4 + e < S && active) ? res[k][(q*4 + e < S) ? q*4 + e : 0] : 0.f;
                                h[e] = to_tf32 (x[e]); l[e] = to_tf32 (x[e] - h[e]);
                                }
                            *reinterpret_cast<float4 *>(base + canon_off (tid, q*4, TM)) = make_float4 (h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<float4 *>(base + A_FLOATS * 4 + canon_off (tid, q*4, TM)) = make_float4 (l[0], l[1], l[2], l[3]);
                            }
                        }
                }
            }
        }

    if (active && sEv.siteDst >= 0)
        ctx.scalers[(size_t)sEv.siteDst * C + c] = site;

    // TMEM no longer needed
    fence_before_sync ();
    __syncthreads ();
    if (warp == 0)
        tmem_dealloc<TMEM_COLS> (tmemBase);

    if (sEv.root < 0)
        return;

    // ---- root integration (Likelihood_Gen, src/likelihood.c:5764-5916), double accumulation ----
    double term = 0.0; int abortFlag = 0;
    if (active)
        {
        const float *rootBase = ctx.partials + (size_t)(sEv.root - ctx.tipCount) * bufStride;
        double like = 0.0;
        for (int k = 0; k < K; k++)
            {
            const float4 *row = reinterpret_cast<const float4 *>(rootBase + ((size_t)k * C + c) * Sp);
            double s = 0.0;
            #pragma unroll
            for (int q = 0; q < NQ; q++)
                {
                const float4 v = __ldcg (row + q);
                s += (double) v.x * freqs[q*4];
                if (q*4 + 1 < S) s += (double) v.y * freqs[q*4 + 1];
                if (q*4 + 2 < S) s += (double) v.z * freqs[q*4 + 2];
                if (q*4 + 3 < S) s += (double) v.w * freqs[q*4 + 3];
                }
            like += s * catW[k];
            }
        double likeI = 0.0;
        if (sEv.hasPInvar)
            {
            const uint64_t im = ctx.invMask[c];
            for (int i = 0; i < S; i++)
                if ((im >> i) & 1) likeI += freqs[i];
            likeI *= sEv.pInvar;
            }
        term = site_term (like, likeI, sEv.hasPInvar, sEv.flags & MB200_QUIRK_FLAG, site,
                          ctx.weights[(size_t)sEv.weightsRow * C + c], abortFlag);
        }
    finish_lnl<128> (ctx, blockIdx.y, term, abortFlag, out, seq);
}

// =================================================================================================
// Node-parallel scheduling (eval_tcq_kernel).  The kernel above walks an evaluation's whole operation list
// inside one CTA per pattern tile: with a few hundred tiles per evaluation (20k codon patterns = 157 tiles
// for 148 SMs) every SM holds about one tile and the ~7 us load -> MMA -> read-out -> store latency of a node
// is paid 30 times back to back.  But the nodes of a tree are only partially ordered: a node needs its two
// children, nothing else.  Here (node, tile) pairs are work items of a device-side queue, handed out level by
// level (height above the clean operands; host-computed), to a persistent grid that fills every SM with as
// many CTAs as fit; an item waits (acquire-polling a flag in global memory) for the items that produce its
// operands, which sit a whole level earlier in the queue and are therefore already running or done.  The
// critical path shrinks from (nodes x latency) to (tree height x latency) and the SMs overlap the phases of
// independent items.  Results are bit-identical to the serial walk: same arithmetic per node, tile partial
// sums still added in tile order.
//   item id -> (slot o, evaluation e, tile t), slot-major; slots 0 .. nOp-1 are the nodes in level order, slot
//   nOp is the evaluation's closing item (site scalers in the caller's operation order, root integration).
//   flags[(e * numTiles + t) * flagStride + node] == seq  <=>  that node's rows of that tile are in memory.
// =================================================================================================
struct TcQueue
{
    unsigned int *counter;      // monotone ticket counter; base = value before this launch's first ticket
    unsigned int  base;
    int          *flags;
    int           flagStride;   // >= most operations of an evaluation + 1
    int           maxOps;       // most operations of any evaluation in this launch
    int           nEval;
    int          *error;        // set when a dependency never arrived (bounded wait)
    const int    *order;        // [total operations] level order, per evaluation at its opOff
};

__device__ __forceinline__ int tcq_ld_acquire (const int *p)
{
    int v;
    asm volatile ("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tcq_st_release (int *p, int v)
{
    asm volatile ("st.release.gpu.global.s32 [%0], %1;\n" :: "l"(p), "r"(v) : "memory");
}

template <int S>
__global__ void __launch_bounds__(128, 1)
eval_tcq_kernel (DevCtx ctx, TcQueue Q, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
                 const DevOp *__restrict__ ops, const float *__restrict__ split, DevResult *out, int seq)
{
    using namespace umma;
    constexpr int NP = TcGeom<S>::NP, KP = TcGeom<S>::KP, KMAX = TcGeom<S>::KMAX;
    constexpr int TM = 128;
    constexpr int ACC_COLS = 2 * NP * KMAX;
    constexpr int TMEM_COLS = (ACC_COLS <= 64) ? 64 : (ACC_COLS <= 128) ? 128 : (ACC_COLS <= 256) ? 256 : 512;
    constexpr uint32_t LBO_A = (TM / 8) * 128, LBO_B = (2 * NP / 8) * 128, SBO = 128;
    constexpr int A_FLOATS = TM * KP;
    constexpr int B_FLOATS = 2 * NP * KP;
    constexpr int NQ = (S + 3) / 4;
    constexpr int SPC = TcGeom<S>::SP;

    extern __shared__ __align__(128) unsigned char tc_smem[];
    float *sA = reinterpret_cast<float *>(tc_smem);           // [k][hi|lo] A images
    float *sB = sA + KMAX * 2 * A_FLOATS;                     // [k] B images (hi, lo)
    float4 *sStage = reinterpret_cast<float4 *>(tc_smem);     // result rows, after the MMAs have completed
    __shared__ uint64_t barB, barM;
    __shared__ uint32_t tmemBase;
    __shared__ int sItem;
    __shared__ double qSum[4];
    __shared__ int    qAb[4];
    __shared__ int    qLast;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sp = ctx.Sp, K = ctx.K, C = ctx.C;
    if (warp == 0)
        tmem_alloc<TMEM_COLS> (&tmemBase);
    const int nIssuers = (2 * K < 4) ? 2 * K : 4;
    if (tid == 0)
        { mbar_init (&barB, 1); mbar_init (&barM, nIssuers); mbar_fence_init (); }
    fence_before_sync ();
    __syncthreads ();
    fence_after_sync ();
    const uint32_t tBase = tmemBase;
    const uint32_t laneSel = (uint32_t)(warp * 32) << 16;
    uint32_t parB = 0, parM = 0;
    const size_t bufStride = (size_t)K * C * Sp;
    const uint64_t fullMask = (S == 64) ? ~(uint64_t)0 : ((((uint64_t)1) << S) - 1);
    constexpr uint32_t idesc = make_idesc_tf32 (TM, NP);
    const uint64_t dA0 = make_desc (smem_u32 (sA), LBO_A, SBO);
    const uint64_t dB0 = make_desc (smem_u32 (sB), LBO_B, SBO);
    constexpr uint64_t A_LO = (A_FLOATS * 4) >> 4, A_K = (2 * A_FLOATS * 4) >> 4, A_KS = (2 * LBO_A) >> 4;
    constexpr uint64_t B_LO = ((NP / 8) * 128) >> 4, B_K = (B_FLOATS * 4) >> 4, B_KS = (2 * LBO_B) >> 4;
    const int rows = ctx.tilePatterns, numTiles = ctx.numTiles;
    const int perSlot = Q.nEval * numTiles;
    const int total = (Q.maxOps + 1) * perSlot;

    for (;;)
        {
        __syncthreads ();                                      // sItem / shared operands of the previous item are free
        if (tid == 0)
            sItem = (int)(atomicAdd (Q.counter, 1u) - Q.base);
        __syncthreads ();
        const int item = sItem;
        if (item >= total || item < 0)
            break;
        const int slot = item / perSlot, e = (item % perSlot) / numTiles, t = item % numTiles;
        const DevEval *ev = evals + e;
        const int nOp = ev->nOp;
        if (slot > nOp)
            continue;
        const int   c0 = t * rows;
        const int   np = min (rows, C - c0);
        const int   c  = c0 + tid;
        const bool  active = tid < np;
        int *flagRow = Q.flags + ((size_t)e * numTiles + t) * Q.flagStride;

        if (slot < nOp)
            {
            // ---------------------------------------------------------------- one node of one tile
            const int   oi = Q.order[ev->opOff + slot];
            const DevOp op = ops[ev->opOff + oi];
            const int   nChild = (op.c3 >= 0) ? 3 : 2;
            const bool  shortcutFlag = (ev->flags & MB200_SHORTCUT_FLAG) != 0;
            if (tid < 3)
                {
                const int pr = (tid == 0) ? op.s1 : (tid == 1) ? op.s2 : op.s3;
                if (pr >= 0)
                    {
                    unsigned spins = 0;
                    while (tcq_ld_acquire (flagRow + pr) != seq)
                        {
                        __nanosleep (100);
                        if ((++spins & 1023u) == 0 && (spins > (1u << 22) || *((volatile int *) Q.error)))
                            { *Q.error = 1; break; }                            // bounded: never hang the device
                        }
                    }
                }
            __syncthreads ();
            float res[KMAX][S];

            for (int ch = 0; ch < nChild; ch++)
                {
                const int child = (ch == 0) ? op.c1 : (ch == 1) ? op.c2 : op.c3;
                const int mat   = (ch == 0) ? op.m1 : (ch == 1) ? op.m2 : op.m3;
                const bool isTip = child < ctx.tipCount;
                bool tipFull = false;
                if (tid == 0)
                    {
                    mbar_expect_tx (&barB, (uint32_t)(K * B_FLOATS * 4));
                    bulk_g2s (sB, split + (size_t)mat * K * B_FLOATS, (uint32_t)(K * B_FLOATS * 4), &barB);
                    }
                if (isTip)
                    {
                    const uint64_t m = active ? ctx.tip64[(size_t)child * C + c] : 0;
                    #pragma unroll
                    for (int q = 0; q < KP / 4; q++)
                        {
                        float4 h;
                        h.x = ((m >> (q*4 + 0)) & 1) ? 1.f : 0.f; h.y = ((m >> (q*4 + 1)) & 1) ? 1.f : 0.f;
                        h.z = ((m >> (q*4 + 2)) & 1) ? 1.f : 0.f; h.w = ((m >> (q*4 + 3)) & 1) ? 1.f : 0.f;
                        *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(sA) + canon_off (tid, q*4, TM)) = h;
                        }
                    tipFull = shortcutFlag && active && m == fullMask && !ctx.tipPartAmbig[child];
                    }
                else
                    {
                    constexpr int QB = (KP / 4 + 3) / 4;
                    constexpr int NIT = (TM / 8) * QB / 4;
                    for (int k = 0; k < K; k++)
                        {
                        const float *src = ctx.partials + (size_t)(child - ctx.tipCount) * bufStride + ((size_t)k * C + c0) * Sp;
                        unsigned char *base = reinterpret_cast<unsigned char *>(sA + (size_t)k * 2 * A_FLOATS);
                        float4 x[NIT];
                        #pragma unroll
                        for (int n = 0; n < NIT; n++)
                            {
                            const int it = warp + 4 * n;
                            const int rb = it / QB, qb = it % QB;
                            const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                            x[n] = make_float4 (0.f, 0.f, 0.f, 0.f);
                            if (q < KP / 4 && r < np && q * 4 < SPC)
                                x[n] = __ldcg (reinterpret_cast<const float4 *>(src + (size_t)r * SPC + q * 4));
                            }
                        #pragma unroll
                        for (int n = 0; n < NIT; n++)
                            {
                            const int it = warp + 4 * n;
                            const int rb = it / QB, qb = it % QB;
                            const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                            if (q < KP / 4)
                                {
                                const float4 h = make_float4 (to_tf32 (x[n].x), to_tf32 (x[n].y), to_tf32 (x[n].z), to_tf32 (x[n].w));
                                const float4 l = make_float4 (to_tf32 (x[n].x - h.x), to_tf32 (x[n].y - h.y), to_tf32 (x[n].z - h.z), to_tf32 (x[n].w - h.w));
                                *reinterpret_cast<float4 *>(base + canon_off (r, q*4, TM)) = h;
                                *reinterpret_cast<float4 *>(base + A_FLOATS * 4 + canon_off (r, q*4, TM)) = l;
                                }
                            }
                        }
                    }
                fence_async_smem ();
                mbar_wait (&barB, parB); parB ^= 1;
                fence_before_sync ();
                __syncthreads ();
                fence_after_sync ();
                if (lane == 0 && warp < nIssuers)
                    {
                    bool any = false;
                    for (int it = warp; it < 2 * K; it += 4)
                        {
                        const int k = it >> 1, corr = it & 1;
                        const uint64_t aHi = dA0 + (isTip ? 0 : (uint64_t)k * A_K), aLo = aHi + A_LO;
                        const uint64_t bHi = dB0 + (uint64_t)k * B_K, bLo = bHi + B_LO;
                        const uint32_t tAcc = tBase + k * 2 * NP + corr * NP;
                        if (!corr)
                            {
                            #pragma unroll
                            for (int ks = 0; ks < KP / 8; ks++)
                                mma_tf32 (tAcc, aHi + ks * A_KS, bHi + ks * B_KS, idesc, ks > 0);
                            }
                        else
                            {
                            #pragma unroll
                            for (int ks = 0; ks < KP / 8; ks++)
                                mma_tf32 (tAcc, aHi + ks * A_KS, bLo + ks * B_KS, idesc, ks > 0);
                            if (!isTip)
                                {
                                #pragma unroll
                                for (int ks = 0; ks < KP / 8; ks++)
                                    mma_tf32 (tAcc, aLo + ks * A_KS, bHi + ks * B_KS, idesc, true);
                                }
                            }
                        any = true;
                        }
                    if (any) mma_commit (&barM);
                    else     mbar_arrive (&barM);
                    }
                mbar_wait (&barM, parM); parM ^= 1;
                fence_after_sync ();
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    {
                    if (k >= K) break;
                    #pragma unroll
                    for (int cb = 0; cb < NP; cb += 32)
                        {
                        uint32_t vm[32], vc[32];
                        tmem_ld32_nowait (tBase + k * 2 * NP + laneSel + cb, vm);
                        tmem_ld32_nowait (tBase + k * 2 * NP + NP + laneSel + cb, vc);
                        tmem_ld_wait ();
                        #pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (cb + i < S)
                                {
                                float v = __uint_as_float (vm[i]) + __uint_as_float (vc[i]);
                                if (tipFull) v = 1.0f;
                                res[k][cb + i] = (ch == 0) ? v : res[k][cb + i] * v;
                                }
                        }
                    }
                fence_before_sync ();
                __syncthreads ();
                fence_after_sync ();
                }

            // rescale (the site-scaler bookkeeping is the closing item's)
            if (active && op.sw >= 0)
                {
                float m = 0.0f;
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    if (k < K)
                        {
                        #pragma unroll
                        for (int i = 0; i < S; i++) m = fmaxf (m, res[k][i]);
                        }
                const float rcp = 1.0f / m;
                #pragma unroll
                for (int k = 0; k < KMAX; k++)
                    if (k < K)
                        {
                        #pragma unroll
                        for (int i = 0; i < S; i++) res[k][i] *= rcp;
                        }
                ctx.scalers[(size_t)op.sw * C + c] = (float) log ((double) m);
                }
            #pragma unroll
            for (int k = 0; k < KMAX; k++)
                if (k < K)
                    {
                    #pragma unroll
                    for (int q = 0; q < NQ; q++)
                        {
                        float4 v;
                        v.x = res[k][q*4];
                        v.y = (q*4 + 1 < S) ? res[k][(q*4 + 1 < S) ? q*4 + 1 : 0] : 0.f;
                        v.z = (q*4 + 2 < S) ? res[k][(q*4 + 2 < S) ? q*4 + 2 : 0] : 0.f;
                        v.w = (q*4 + 3 < S) ? res[k][(q*4 + 3 < S) ? q*4 + 3 : 0] : 0.f;
                        sStage[((size_t)k * NQ + q) * (TM + 1) + tid] = v;
                        }
                    }
            __syncthreads ();
            {
            float *dstBase = ctx.partials + (size_t)(op.dest - ctx.tipCount) * bufStride;
            constexpr int nq = SPC / 4;
            for (int k = 0; k < K; k++)
                {
                float4 *dst = reinterpret_cast<float4 *>(dstBase + ((size_t)k * C + c0) * SPC);
                #pragma unroll
                for (int n = 0; n < nq; n++)
                    {
                    const int idx = n * 128 + tid;
                    const int r = idx / nq, q = idx % nq;
                    if (r < np)
                        dst[idx] = (q < NQ) ? sStage[((size_t)k * NQ + q) * (TM + 1) + r] : make_float4 (0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            // publish: every thread's stores are ordered before the flag (fence, barrier, release store)
            __threadfence ();
            __syncthreads ();
            if (tid == 0)
                tcq_st_release (flagRow + oi, seq);
            continue;
            }

        // -------------------------------------------------------------------- closing item of (e, t)
        for (int o = tid; o < nOp; o += 128)
            {
            unsigned spins = 0;
            while (tcq_ld_acquire (flagRow + o) != seq)
                {
                __nanosleep (200);
                if ((++spins & 1023u) == 0 && (spins > (1u << 22) || *((volatile int *) Q.error)))
                    { *Q.error = 1; break; }
                }
            }
        __syncthreads ();
        float site = 0.0f;
        if (active)
            {
            site = (ev->siteSrc >= 0) ? ctx.scalers[(size_t)ev->siteSrc * C + c] : 0.0f;
            for (int o = 0; o < nOp; o++)                      // the caller's operation order (src/likelihood.c:7938-7965)
                {
                const DevOp op = ops[ev->opOff + o];
                if (op.sr >= 0) site -= __ldcg (ctx.scalers + (size_t)op.sr * C + c);
                if (op.sw >= 0) site += __ldcg (ctx.scalers + (size_t)op.sw * C + c);
                }
            if (ev->siteDst >= 0)
                ctx.scalers[(size_t)ev->siteDst * C + c] = site;
            }
        if (ev->root < 0)
            continue;
        double term = 0.0; int abortFlag = 0;
        if (active)
            {
            const double *catW = dvals + ev->dOff + K, *freqs = dvals + ev->dOff + 2*K;
            const float *rootBase = ctx.partials + (size_t)(ev->root - ctx.tipCount) * bufStride;
            double like = 0.0;
            for (int k = 0; k < K; k++)
                {
                const float4 *row = reinterpret_cast<const float4 *>(rootBase + ((size_t)k * C + c) * Sp);
                double s = 0.0;
                #pragma unroll
                for (int q = 0; q < NQ; q++)
                    {
                    const float4 v = __ldcg (row + q);
                    s += (double) v.x * freqs[q*4];
                    if (q*4 + 1 < S) s += (double) v.y * freqs[q*4 + 1];
                    if (q*4 + 2 < S) s += (double) v.z * freqs[q*4 + 2];
                    if (q*4 + 3 < S) s += (double) v.w * freqs[q*4 + 3];
                    }
                like += s * catW[k];
                }
            double likeI = 0.0;
            if (ev->hasPInvar)
                {
                const uint64_t im = ctx.invMask[c];
                for (int i = 0; i < S; i++)
                    if ((im >> i) & 1) likeI += freqs[i];
                likeI *= ev->pInvar;
                }
            term = site_term (like, likeI, ev->hasPInvar, ev->flags & MB200_QUIRK_FLAG, site,
                              ctx.weights[(size_t)ev->weightsRow * C + c], abortFlag);
            }
        // tile partial -> ticket -> the last tile of the evaluation adds the partials in tile order
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            {
            term      += __shfl_xor_sync (0xffffffffu, term, off);
            abortFlag |= __shfl_xor_sync (0xffffffffu, abortFlag, off);
            }
        if (lane == 0) { qSum[warp] = term; qAb[warp] = abortFlag; }
        __syncthreads ();
        if (tid == 0)
            {
            const double s = qSum[0] + qSum[1] + qSum[2] + qSum[3];
            const int    a = qAb[0] | qAb[1] | qAb[2] | qAb[3];
            ctx.tilePartial[(size_t)e * numTiles + t] = s;
            ctx.tileAbort  [(size_t)e * numTiles + t] = a;
            __threadfence ();
            const unsigned int tk = atomicAdd (&ctx.ticket[e], 1u);
            qLast = (tk == (unsigned int) numTiles - 1u);
            if (qLast)
                {
                __threadfence ();
                double tot = 0.0; int ab = 0;
                for (int tIdx = 0; tIdx < numTiles; tIdx++)
                    {
                    tot += __ldcg (&ctx.tilePartial[(size_t)e * numTiles + tIdx]);
                    ab  |= __ldcg (&ctx.tileAbort  [(size_t)e * numTiles + tIdx]);
                    }
                if (*((volatile int *) Q.error)) ab = 1;
                const double lnL = ab ? -DBL_MAX : tot;
                int4 pkt;
                pkt.x = __double2loint (lnL); pkt.y = __double2hiint (lnL); pkt.z = ab ? 1 : 0; pkt.w = seq;
                *reinterpret_cast<int4 *>(&out[e]) = pkt;
                ctx.ticket[e] = 0u;
                }
            }
        }

    fence_before_sync ();
    __syncthreads ();
    if (warp == 0)
        tmem_dealloc<TMEM_COLS> (tmemBase);
}
