// mb200_kernels_tcp.cuh -- warp-specialised, pipelined tcgen05 pruning kernel (eval_tcp_kernel) for the 20-state
// amino-acid and 61-state codon paths (CondLikeDown/Root_Gen*, _NY98*, CondLikeScaler_Gen*, Likelihood_Gen*/_NY98*;
// reference src/likelihood.c:204, 1575, 2152, 4010, 4939, 5413, 5764, 6975).
//
// Work = (node, 128-pattern tile) items of the device-side queue of mb200_kernels_tc.cuh (level order, acquire /
// release flags between the CTAs of a persistent grid).  Inside the CTA the phases of an item do not run back to
// back: one CTA per SM, 20 warps with fixed roles, mbarrier rings between them.
//
//   scheduler (1 warp)    draws tickets, decodes (slot, evaluation, tile), waits for the item's producers (flag
//                         acquire), publishes the item in a 4-deep shared-memory ring
//   loaders   (8 warps)   one *unit* = one (rate category, child) operand of an item.  Groups of warps take the units
//                         round-robin, so several units' loads are in flight per SM: child rows HBM/L2 -> registers
//                         (512-byte coalesced LDG.128) -> hi/lo TF32 split -> canonical K-major core-matrix images in
//                         an operand-ring stage; the branch's pre-split P(t) image [hi | lo] arrives in the same stage by
//                         one bulk async copy (TMA engine, complete_tx on the stage's full barrier)
//   MMA       (1 warp)    per unit: [main | corr] = A_hi x [B_hi | B_lo]^T  (ONE tcgen05.mma chain, N = 2 NP) and
//                         corr += A_lo x B_hi^T (N = NP) into one slot of a TMEM accumulator ring; tcgen05.commit
//                         frees the operand stage and hands the accumulators to the epilogue
//   epilogue  (2 x 4)     two halves share an item: thread = pattern row = TMEM lane, the 16-column strips of a row
//                         alternate between the halves: tcgen05.ld, main + corr, product over the children, row maximum;
//                         unscaled rows staged in shared memory, then coalesced 16-byte stores of row * (1 / max) -- the
//                         two roundings of the reference's rescaler; node scaler; the evaluation's closing item (site
//                         scalers, root integration, lnL tile sums) runs here too
//   publisher (1 warp)    release-stores the node-done flags (the memory barrier of a release does not stall a warp that
//                         has accumulators waiting)
//
// 3xTF32 as before (x = hi + lo, lo x lo dropped); the large term and the two correction terms land in separate
// accumulators and are added in FP32 in the epilogue.
#pragma once
#include "mb200_kernels_tc.cuh"

template <int S> struct TcpGeom;
// NA: TMEM ring, NA units x 2 NP columns = 512.  LG: loader groups (8 / LG warps each) -- units go round-robin to the
// groups, so LG units' loads are in flight per SM; a 20-state unit is small (10 KB), hence more, smaller groups
template <> struct TcpGeom<61> { static constexpr int NA = 4, LG = 2; };
template <> struct TcpGeom<20> { static constexpr int NA = 8, LG = 4; };

constexpr int TCP_THREADS   = 608;     // 2 x 4 epilogue warps, 8 loader warps, MMA issuer, scheduler, publisher
constexpr int TCP_NPUB      = 4;       // flag-publication ring (epilogue -> publisher warp)
constexpr int TCP_NS_MAX    = 8;       // operand-ring stages (as many as fit beside the staging area)
constexpr int TCP_NI        = 4;       // item ring
constexpr int TCP_ITEM_NODE = 0, TCP_ITEM_CLOSE = 1, TCP_ITEM_STOP = 2;

struct TcpItem                          // 64 bytes
{
    int kind, e, t, oi;
    int nChild, dest, sw, shortcut;
    int child[3], mat[3];
    int pad[2];
};

template <int S> __host__ __device__ constexpr size_t tcp_stage_bytes ()
{
    return (size_t)(2 * 128 * TcGeom<S>::KP + 2 * TcGeom<S>::NP * TcGeom<S>::KP) * sizeof(float);
}
template <int S> __host__ __device__ constexpr size_t tcp_staging_bytes (int K)
{
    return (size_t) K * ((S + 3) / 4) * 129 * sizeof(float4);   // one item's result rows, [k][16-byte chunk][row + pad]
}
// "this row's tip is fully ambiguous" bytes, one 128-byte record per unit in flight (loader -> epilogue)
template <int S> __host__ __device__ constexpr size_t tcp_tipring_bytes (int NS) { return (size_t)(TcpGeom<S>::NA + NS) * 128; }
// stages that fit in `limit` bytes of dynamic shared memory: a multiple of the loader groups (every stage is then
// always filled by the same group, which sees each of its phases -- the parity wait cannot alias), or 1: one group
// loads everything
template <int S> inline int tcp_stages (int K, size_t limit)
{
    const size_t st = tcp_staging_bytes<S> (K) + tcp_tipring_bytes<S> (0);
    if (st + tcp_stage_bytes<S> () + 128 > limit) return 0;
    size_t n = (limit - st) / (tcp_stage_bytes<S> () + 128);
    if (n > (size_t) TCP_NS_MAX) n = TCP_NS_MAX;
    if (n >= (size_t) TcpGeom<S>::LG) n -= n % TcpGeom<S>::LG;
    else if (n >= 2) n = 2;
    return (int) n;
}

#ifndef TCP_BACKOFF_NS
#define TCP_BACKOFF_NS 32
#endif
// mbarrier wait that cannot hang the device: a CTA whose pipeline stalls for two seconds traps (the launch fails
// with an error instead of sitting on the GPU until somebody's watchdog fires)
__device__ __forceinline__ void tcp_wait (uint64_t *bar, uint32_t parity)
{
    const uint32_t a = umma::smem_u32 (bar);
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (unsigned it = 0; ; it++)
        {
        asm volatile ("{\n\t.reg .pred p;\n\t"
                      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                      "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
#ifndef TCP_NO_BACKOFF
        __nanosleep (TCP_BACKOFF_NS);          // a waiting warp must not compete for issue slots with the working ones
#endif
        if ((it & 0xfffu) == 0xfffu)
            {
            unsigned long long now;
            asm volatile ("mov.u64 %0, %%globaltimer;\n" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap ();
            }
        }
}
__device__ __forceinline__ void tcp_bar_epilogue () { asm volatile ("bar.sync 3, 256;\n" ::: "memory"); }      // both epilogue halves
// barrier of one epilogue half (warps 0-3: id 1, warps 4-7: id 2)
__device__ __forceinline__ void tcp_bar_group (int grp)
{
    if (grp == 0) asm volatile ("bar.sync 1, 128;\n" ::: "memory");
    else          asm volatile ("bar.sync 2, 128;\n" ::: "memory");
}

// debug builds (-DMB200_PHASE_TIMING): CTA 0 stamps the pipeline events of its first items, 16 slots per item
#ifdef MB200_PHASE_TIMING
#define TCP_TRACE_ITEMS 250
#define TCP_T(n, slot) do { if (blockIdx.x == 0 && (n) < TCP_TRACE_ITEMS) ctx.dbg[(n) * 16 + (slot)] = mb200_now (); } while (0)
#define TCP_TV(n, slot, v) do { if (blockIdx.x == 0 && (n) < TCP_TRACE_ITEMS) ctx.dbg[(n) * 16 + (slot)] = (unsigned long long)(v); } while (0)
#else
#define TCP_T(n, slot) do { } while (0)
#define TCP_TV(n, slot, v) do { } while (0)
#endif

template <int S>
__global__ void __launch_bounds__(TCP_THREADS, 1)
eval_tcp_kernel (DevCtx ctx, TcQueue Q, int NS, const DevEval *__restrict__ evals, const double *__restrict__ dvals,
                 const DevOp *__restrict__ ops, const float *__restrict__ split, DevResult *out, int seq)
{
    using namespace umma;
    constexpr int NP = TcGeom<S>::NP, KP = TcGeom<S>::KP, NA = TcpGeom<S>::NA;
    constexpr int TM = 128;
    constexpr int UC = 2 * NP;                                  // TMEM columns of one unit: [main | corr]
    constexpr uint32_t LBO_A = (TM / 8) * 128, LBO_B = (2 * NP / 8) * 128, SBO = 128;
    constexpr int A_FLOATS = TM * KP;
    constexpr int B_FLOATS = 2 * NP * KP;
    constexpr int NQ = (S + 3) / 4;
    constexpr int SPC = TcGeom<S>::SP;
    constexpr size_t STAGE = tcp_stage_bytes<S> ();

    extern __shared__ __align__(128) unsigned char tcp_smem[];
    __shared__ uint64_t barFull[TCP_NS_MAX], barEmpty[TCP_NS_MAX], barAccFull[NA], barAccEmpty[NA], barInfoFull[TCP_NI], barInfoEmpty[TCP_NI];
    __shared__ TcpItem sInfo[TCP_NI];
    __shared__ uint32_t tmemBase;
    __shared__ uint64_t barPubFull[TCP_NPUB], barPubEmpty[TCP_NPUB];
    __shared__ int   *sPub[TCP_NPUB];  // flags to release, in order (nullptr: stop)
    __shared__ float  sMax[2][TM];     // row maxima found by the two epilogue halves
    __shared__ double qSum[2][4];
    __shared__ int    qAb[2][4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sp = ctx.Sp, K = ctx.K, C = ctx.C;
    float4 *sStage = reinterpret_cast<float4 *>(tcp_smem + (size_t) NS * STAGE);     // [group][buffer][chunk][TM+1]
    const int TIPRING = NA + NS;
    unsigned char *sTipFull = tcp_smem + (size_t) NS * STAGE + tcp_staging_bytes<S> (K);   // [unit % TIPRING][row]

    if (warp == 0)
        tmem_alloc<512> (&tmemBase);
    if (tid == 32)
        {
        for (int s = 0; s < NS; s++) { mbar_init (&barFull[s], 8 / TcpGeom<S>::LG + 1); mbar_init (&barEmpty[s], 1); }
        for (int a = 0; a < NA; a++) { mbar_init (&barAccFull[a], 2); mbar_init (&barAccEmpty[a], 8); }
        for (int i = 0; i < TCP_NI; i++) { mbar_init (&barInfoFull[i], 1); mbar_init (&barInfoEmpty[i], 17); }
        for (int i = 0; i < TCP_NPUB; i++) { mbar_init (&barPubFull[i], 1); mbar_init (&barPubEmpty[i], 1); }
        mbar_fence_init ();
        }
    fence_before_sync ();
    __syncthreads ();
    fence_after_sync ();
    const uint32_t tBase = tmemBase;
    const size_t bufStride = (size_t)K * C * Sp;
    const int rows = TM, numTiles = ctx.numTiles;
    const int perSlot = Q.nEval * numTiles;
    const int total = (Q.maxOps + 1) * perSlot;

    if (warp == 17)
        {
        // =================================================================== scheduler
        int slot = 0; uint32_t ph = 0;
        unsigned nItem = 0;
        for (;;)
            {
            tcp_wait (&barInfoEmpty[slot], ph ^ 1);
            if (lane == 0) TCP_T (nItem, 1);
            int item = 0;
            if (lane == 0) item = (int)(atomicAdd (Q.counter, 1u) - Q.base);
            item = __shfl_sync (0xffffffffu, item, 0);
            TcpItem it;
            it.kind = TCP_ITEM_STOP; it.e = it.t = it.oi = 0; it.nChild = 0; it.dest = it.sw = -1; it.shortcut = 0;
            it.child[0] = it.child[1] = it.child[2] = -1; it.mat[0] = it.mat[1] = it.mat[2] = -1; it.pad[0] = it.pad[1] = 0;
            if (item >= 0 && item < total)
                {
                const int sl = item / perSlot, e = (item % perSlot) / numTiles, t = item % numTiles;
                const DevEval *ev = evals + e;
                const int nOp = ev->nOp;
                if (sl > nOp)
                    continue;                                   // shorter evaluation: nothing in this slot
                const int *flagRow = Q.flags + ((size_t)e * numTiles + t) * Q.flagStride;
                it.e = e; it.t = t;
                if (sl < nOp)
                    {
                    const int   oi = Q.order[ev->opOff + sl];
                    const DevOp op = ops[ev->opOff + oi];
                    it.kind = TCP_ITEM_NODE; it.oi = oi;
                    it.nChild = (op.c3 >= 0) ? 3 : 2; it.dest = op.dest; it.sw = op.sw;
                    it.shortcut = (ev->flags & MB200_SHORTCUT_FLAG) ? 1 : 0;
                    it.child[0] = op.c1; it.child[1] = op.c2; it.child[2] = op.c3;
                    it.mat[0] = op.m1; it.mat[1] = op.m2; it.mat[2] = op.m3;
                    if (lane == 0) TCP_T (nItem, 2);
                    if (lane < 3)
                        {
                        const int pr = (lane == 0) ? op.s1 : (lane == 1) ? op.s2 : op.s3;
                        if (pr >= 0)
                            {
                            unsigned spins = 0;
                            while (tcq_ld_acquire (flagRow + pr) != seq)
                                {
                                __nanosleep (64);
                                if ((++spins & 1023u) == 0 && (spins > (1u << 22) || *((volatile int *) Q.error)))
                                    { *Q.error = 1; break; }                    // bounded: never hang the device
                                }
                            }
                        }
                    }
                else
                    {
                    it.kind = TCP_ITEM_CLOSE;
                    for (int o = lane; o < nOp; o += 32)
                        {
                        unsigned spins = 0;
                        while (tcq_ld_acquire (flagRow + o) != seq)
                            {
                            __nanosleep (128);
                            if ((++spins & 1023u) == 0 && (spins > (1u << 22) || *((volatile int *) Q.error)))
                                { *Q.error = 1; break; }
                            }
                        }
                    }
                __syncwarp ();
                }
            if (lane == 0)
                {
                sInfo[slot] = it;
                mbar_arrive (&barInfoFull[slot]);              // release: the record is visible to whoever acquires the phase
                TCP_TV (nItem, 0, (unsigned long long) it.kind | ((unsigned long long) it.t << 8) | ((unsigned long long) it.oi << 32));
                TCP_T (nItem, 3);
                }
            nItem++;
            if (it.kind == TCP_ITEM_STOP)
                break;
            if (++slot == TCP_NI) { slot = 0; ph ^= 1; }
            }
        }
    else if (warp >= 8 && warp < 16)
        {
        // =================================================================== loaders: groups of warps, units round-robin
        constexpr int LGMAX = TcpGeom<S>::LG, WPG = 8 / LGMAX;     // warps per group
        const int nGroups = (NS >= LGMAX) ? LGMAX : (NS >= 2) ? 2 : 1;          // NS is a multiple of nGroups
        const int grp = (warp - 8) / WPG, lw = (warp - 8) % WPG, ltid = lw * 32 + lane;
        constexpr int QB = (KP / 4 + 3) / 4;                    // chunk blocks of 4 per row
        constexpr int NIT = (TM / 8) * QB / WPG;                // items per thread and unit
        constexpr int TPR = TM / (WPG * 32);                    // tip rows per thread
        const uint64_t fullMaskL = (S == 64) ? ~(uint64_t)0 : ((((uint64_t)1) << S) - 1);
        int islot = 0; uint32_t iph = 0;
        unsigned u = 0;                                         // units since the kernel started (all roles count alike)
        unsigned nItem = 0;
        for (;;)
            {
            tcp_wait (&barInfoFull[islot], iph);
            const int kind = sInfo[islot].kind, tileIdx = sInfo[islot].t, nChild = sInfo[islot].nChild;
            const int ch0 = sInfo[islot].child[0], ch1 = sInfo[islot].child[1], ch2 = sInfo[islot].child[2];
            const int mt0 = sInfo[islot].mat[0], mt1 = sInfo[islot].mat[1], mt2 = sInfo[islot].mat[2];
            __syncwarp ();
            if (lane == 0) mbar_arrive (&barInfoEmpty[islot]);
            if (++islot == TCP_NI) { islot = 0; iph ^= 1; }
            if (kind == TCP_ITEM_STOP) break;
            if (kind == TCP_ITEM_CLOSE) { nItem++; continue; }
            const int c0 = tileIdx * rows, np = min (rows, C - c0);
            bool firstUnit = true;
            for (int k = 0; k < K; k++)
                for (int ch = 0; ch < nChild; ch++, u++)
                    {
                    if ((int)(u % (unsigned) nGroups) != grp)
                        continue;
                    const int s = (int)(u % (unsigned) NS);
                    const uint32_t ph = (u / (unsigned) NS) & 1u;
                    const int child = (ch == 0) ? ch0 : (ch == 1) ? ch1 : ch2, mat = (ch == 0) ? mt0 : (ch == 1) ? mt1 : mt2;
                    const bool isTip = child < ctx.tipCount;
                    unsigned char *stg = tcp_smem + (size_t) s * STAGE;
                    float4 x[NIT];
                    uint64_t m[TPR];
                    int partAmbig = 0;
                    // loads first, then the wait for the stage: the round trip overlaps the MMAs still reading it
                    if (isTip)
                        {
                        partAmbig = ctx.tipPartAmbig[child];
                        #pragma unroll
                        for (int j = 0; j < TPR; j++)
                            {
                            const int r = ltid + j * (WPG * 32);
                            m[j] = (r < np) ? ctx.tip64[(size_t)child * C + c0 + r] : 0;
                            }
                        }
                    else
                        {
                        const float *src = ctx.partials + (size_t)(child - ctx.tipCount) * bufStride + ((size_t)k * C + c0) * Sp;
                        #pragma unroll
                        for (int n = 0; n < NIT; n++)
                            {
                            const int i2 = lw + WPG * n;
                            const int rb = i2 / QB, qb = i2 % QB;
                            const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                            x[n] = make_float4 (0.f, 0.f, 0.f, 0.f);
                            if (q < KP / 4 && r < np && q * 4 < SPC)
                                x[n] = __ldcg (reinterpret_cast<const float4 *>(src + (size_t)r * SPC + q * 4));
                            }
                        }
                    tcp_wait (&barEmpty[s], ph ^ 1);
#ifndef TCP_TRACE_EPI
                    if (ltid == 0 && firstUnit && grp < 2) TCP_T (nItem, 4 + 2 * grp);
#endif
                    if (ltid == 0)
                        {
                        mbar_expect_tx (&barFull[s], (uint32_t)(B_FLOATS * 4));
                        bulk_g2s (stg + 2 * A_FLOATS * 4, split + ((size_t)mat * K + k) * B_FLOATS, (uint32_t)(B_FLOATS * 4), &barFull[s]);
                        }
                    if (isTip)
                        {
                        // thread = row(s): the state mask expands to 0/1 (exact in TF32; no lo image); the epilogue
                        // learns through the tip ring which rows see a fully ambiguous tip (preLike shortcut)
                        #pragma unroll
                        for (int j = 0; j < TPR; j++)
                            {
                            const int r = ltid + j * (WPG * 32);
                            sTipFull[(u % (unsigned) TIPRING) * 128 + r] = (m[j] == fullMaskL && !partAmbig) ? 1 : 0;
                            #pragma unroll
                            for (int q = 0; q < KP / 4; q++)
                                {
                                float4 h;
                                h.x = ((m[j] >> (q*4 + 0)) & 1) ? 1.f : 0.f; h.y = ((m[j] >> (q*4 + 1)) & 1) ? 1.f : 0.f;
                                h.z = ((m[j] >> (q*4 + 2)) & 1) ? 1.f : 0.f; h.w = ((m[j] >> (q*4 + 3)) & 1) ? 1.f : 0.f;
                                *reinterpret_cast<float4 *>(stg + canon_off (r, q*4, TM)) = h;
                                }
                            }
                        }
                    else
                        {
                        #pragma unroll
                        for (int n = 0; n < NIT; n++)
                            {
                            const int i2 = lw + WPG * n;
                            const int rb = i2 / QB, qb = i2 % QB;
                            const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
                            if (q < KP / 4)
                                {
                                const float4 h = make_float4 (to_tf32 (x[n].x), to_tf32 (x[n].y), to_tf32 (x[n].z), to_tf32 (x[n].w));
                                const float4 l = make_float4 (to_tf32 (x[n].x - h.x), to_tf32 (x[n].y - h.y), to_tf32 (x[n].z - h.z), to_tf32 (x[n].w - h.w));
                                *reinterpret_cast<float4 *>(stg + canon_off (r, q*4, TM)) = h;
                                *reinterpret_cast<float4 *>(stg + A_FLOATS * 4 + canon_off (r, q*4, TM)) = l;
                                }
                            }
                        }
                    fence_async_smem ();                        // generic-proxy stores -> async proxy (tcgen05.mma)
                    __syncwarp ();
                    if (lane == 0) mbar_arrive (&barFull[s]);
#ifndef TCP_TRACE_EPI
                    if (ltid == 0 && firstUnit && grp < 2) TCP_T (nItem, 5 + 2 * grp);
#endif
                    firstUnit = false;
                    }
            nItem++;
            }
        }
    else if (warp == 16)
        {
        // =================================================================== MMA issuer
        constexpr uint32_t idescWide = make_idesc_tf32 (TM, 2 * NP), idescNarrow = make_idesc_tf32 (TM, NP);
        constexpr uint64_t A_LO = (A_FLOATS * 4) >> 4, A_KS = (2 * LBO_A) >> 4, B_KS = (2 * LBO_B) >> 4, ST = STAGE >> 4;
        const uint64_t dA0 = make_desc (smem_u32 (tcp_smem), LBO_A, SBO);
        const uint64_t dB0 = make_desc (smem_u32 (tcp_smem + 2 * A_FLOATS * 4), LBO_B, SBO);
        int islot = 0; uint32_t iph = 0;
        unsigned u = 0;
        unsigned nItem = 0;
        for (;;)
            {
            tcp_wait (&barInfoFull[islot], iph);
            const int kind = sInfo[islot].kind, nChild = sInfo[islot].nChild;
            const int c1 = sInfo[islot].child[0], c2 = sInfo[islot].child[1], c3 = sInfo[islot].child[2];
            __syncwarp ();
            if (lane == 0) mbar_arrive (&barInfoEmpty[islot]);
            if (++islot == TCP_NI) { islot = 0; iph ^= 1; }
            if (kind == TCP_ITEM_STOP) break;
            if (kind == TCP_ITEM_CLOSE) { nItem++; continue; }
            for (int k = 0; k < K; k++)
                for (int ch = 0; ch < nChild; ch++, u++)
                    {
                    const int s = (int)(u % (unsigned) NS), a = (int)(u % (unsigned) NA);
                    const uint32_t ph = (u / (unsigned) NS) & 1u, aph = (u / (unsigned) NA) & 1u;
                    const bool isTip = ((ch == 0) ? c1 : (ch == 1) ? c2 : c3) < ctx.tipCount;
                    tcp_wait (&barFull[s], ph);                 // operand images (generic stores + bulk copy) have landed
                    if (lane == 0 && k == 0 && ch == 0) TCP_T (nItem, 8);
                    tcp_wait (&barAccEmpty[a], aph ^ 1);        // the epilogue has drained this accumulator slot
                    if (lane == 0 && k == 0 && ch == 0) TCP_T (nItem, 9);
                    fence_after_sync ();
                    if (lane == 0)
                        {
                        const uint64_t aHi = dA0 + (uint64_t) s * ST, aLo = aHi + A_LO, bb = dB0 + (uint64_t) s * ST;
                        const uint32_t tAcc = tBase + (uint32_t)(a * UC);
                        #pragma unroll
                        for (int ks = 0; ks < KP / 8; ks++)
                            mma_tf32 (tAcc, aHi + ks * A_KS, bb + ks * B_KS, idescWide, ks > 0);
                        if (!isTip)
                            {
                            #pragma unroll
                            for (int ks = 0; ks < KP / 8; ks++)
                                mma_tf32 (tAcc + NP, aLo + ks * A_KS, bb + ks * B_KS, idescNarrow, true);
                            }
                        mma_commit (&barEmpty[s]);              // operand stage free once these MMAs have read it
                        mma_commit (&barAccFull[a]);            // accumulators complete ...
                        mbar_arrive (&barAccFull[a]);           // ... and what this thread has seen (the loaders' shared-memory
                                                                // writes, acquired with the stage) is released to the epilogue
                        if (k == K - 1 && ch == nChild - 1) TCP_T (nItem, 10);
                        }
                    __syncwarp ();
                    }
            nItem++;
            }
        }
    else if (warp == 18)
        {
        // =================================================================== publisher: release-stores the node-done flags,
        // so that no epilogue warp sits in a memory barrier
        int slot = 0; uint32_t ph = 0;
        for (;;)
            {
            tcp_wait (&barPubFull[slot], ph);
            int *flag = sPub[slot];
            __syncwarp ();
            if (lane == 0)
                {
                mbar_arrive (&barPubEmpty[slot]);
                if (flag != nullptr)
                    tcq_st_release (flag, seq);                // fence + store: everything the epilogue wrote is visible first
                }
            if (flag == nullptr)
                break;
            if (++slot == TCP_NPUB) { slot = 0; ph ^= 1; }
            }
        }
    else
        {
        // =================================================================== epilogue: two halves of four warps share every
        // item -- thread = pattern row = TMEM lane in both, the 16-column strips of a row alternate between them; the
        // products are staged UNSCALED in shared memory ([k][chunk][row]: conflict-free for thread = row), the halves
        // exchange their row maxima, and all 256 threads copy the rows out with coalesced 16-byte stores of
        // row * (1 / max) -- the two roundings of CondLikeScaler_Gen (src/likelihood.c:4939-4990).
        const int grp = warp >> 2, gtid = tid & (TM - 1), row = gtid;
        const uint32_t laneSel = (uint32_t)((warp & 3) * 32) << 16;
        int islot = 0; uint32_t iph = 0;
        int pslot = 0; uint32_t pph = 0;
        unsigned u = 0;
        unsigned nItem = 0;
        unsigned accPar = 0;                                    // parity of the next phase of barAccFull[slot], one bit per slot
        for (;; nItem++)
            {
            tcp_wait (&barInfoFull[islot], iph);
            if (tid == 0) TCP_T (nItem, 11);
            struct { int kind, e, t, oi, nChild, dest, sw, shortcut; } it;
            it.kind = sInfo[islot].kind; it.e = sInfo[islot].e; it.t = sInfo[islot].t; it.oi = sInfo[islot].oi;
            it.nChild = sInfo[islot].nChild; it.dest = sInfo[islot].dest; it.sw = sInfo[islot].sw; it.shortcut = sInfo[islot].shortcut;
            const int ch0 = sInfo[islot].child[0], ch1 = sInfo[islot].child[1], ch2 = sInfo[islot].child[2];
            __syncwarp ();
            if (lane == 0) mbar_arrive (&barInfoEmpty[islot]);
            if (++islot == TCP_NI) { islot = 0; iph ^= 1; }
            if (it.kind == TCP_ITEM_STOP)
                {
                if (tid == 0)
                    {
                    tcp_wait (&barPubEmpty[pslot], pph ^ 1);
                    sPub[pslot] = nullptr;
                    mbar_arrive (&barPubFull[pslot]);
                    }
                break;
                }
            const int   c0 = it.t * rows, np = min (rows, C - c0);
            const int   c  = c0 + row;
            const bool  active = row < np;
            const DevEval *ev = evals + it.e;
            int *flagRow = Q.flags + ((size_t)it.e * numTiles + it.t) * Q.flagStride;

            if (it.kind == TCP_ITEM_NODE)
                {
                float mx = 0.0f;
                // preLike shortcut of the scalar kernels (src/likelihood.c:257-258): a fully ambiguous tip contributes exactly 1
                const unsigned tipKids = it.shortcut ? ((ch0 < ctx.tipCount ? 1u : 0u) | (ch1 < ctx.tipCount ? 2u : 0u) |
                                                        ((it.nChild > 2 && ch2 < ctx.tipCount) ? 4u : 0u)) : 0u;
                for (int k = 0; k < K; k++)
                    {
                    // the accumulators of all children of (item, k) sit in consecutive slots of the TMEM ring: wait for
                    // the last one (commits complete in issue order), then combine 16 columns at a time
                    for (int ch = 0; ch < it.nChild; ch++)
                        {
                        const int a = (int)((u + ch) % (unsigned) NA);
                        tcp_wait (&barAccFull[a], (accPar >> a) & 1u);
                        accPar ^= 1u << a;
                        }
                    fence_after_sync ();
                    if (gtid == 0 && k == K - 1) TCP_T (nItem, 12);
                    unsigned tipFull = 0;
                    if (tipKids)
                        {
                        if (tipKids & 1u) tipFull |= sTipFull[(u % (unsigned) TIPRING) * 128 + row] ? 1u : 0u;
                        if (tipKids & 2u) tipFull |= sTipFull[((u + 1) % (unsigned) TIPRING) * 128 + row] ? 2u : 0u;
                        if (tipKids & 4u) tipFull |= sTipFull[((u + 2) % (unsigned) TIPRING) * 128 + row] ? 4u : 0u;
                        }
                    const uint32_t tA0 = tBase + (uint32_t)((int)(u % (unsigned) NA) * UC) + laneSel;
                    const uint32_t tA1 = tBase + (uint32_t)((int)((u + 1) % (unsigned) NA) * UC) + laneSel;
                    const uint32_t tA2 = tBase + (uint32_t)((int)((u + 2) % (unsigned) NA) * UC) + laneSel;
                    #pragma unroll
                    for (int cb = 0; cb < NP; cb += 16)
                        {
                        if (cb >= S) break;
                        if ((((cb >> 4) + k) & 1) != grp) continue;     // the other half's strip
                        float prod[16];
                        {
                        // the first two children's strips are read together: four TMEM loads in flight, one wait
                        uint32_t vm0[16], vc0[16], vm1[16], vc1[16];
                        tmem_ld16_nowait (tA0 + cb, vm0);
                        tmem_ld16_nowait (tA0 + NP + cb, vc0);
                        tmem_ld16_nowait (tA1 + cb, vm1);
                        tmem_ld16_nowait (tA1 + NP + cb, vc1);
                        tmem_ld_wait ();
                        #pragma unroll
                        for (int i = 0; i < 16; i++)
                            {
                            float v0 = __uint_as_float (vm0[i]) + __uint_as_float (vc0[i]);
                            float v1 = __uint_as_float (vm1[i]) + __uint_as_float (vc1[i]);
                            if (tipFull & 1u) v0 = 1.0f;
                            if (tipFull & 2u) v1 = 1.0f;
                            prod[i] = v0 * v1;
                            }
                        }
                        if (it.nChild > 2)
                            {
                            uint32_t vm[16], vc[16];
                            tmem_ld16_nowait (tA2 + cb, vm);
                            tmem_ld16_nowait (tA2 + NP + cb, vc);
                            tmem_ld_wait ();
                            #pragma unroll
                            for (int i = 0; i < 16; i++)
                                {
                                float v = __uint_as_float (vm[i]) + __uint_as_float (vc[i]);
                                if (tipFull & 4u) v = 1.0f;
                                prod[i] *= v;
                                }
                            }
                        #pragma unroll
                        for (int i = 0; i < 16; i++)
                            if (cb + i < S) mx = fmaxf (mx, prod[i]);
                        #pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (cb + j*4 < S)
                                {
                                float4 v;
                                v.x = prod[j*4];
                                v.y = (cb + j*4 + 1 < S) ? prod[j*4 + 1] : 0.f;
                                v.z = (cb + j*4 + 2 < S) ? prod[j*4 + 2] : 0.f;
                                v.w = (cb + j*4 + 3 < S) ? prod[j*4 + 3] : 0.f;
                                sStage[((size_t)k * NQ + cb / 4 + j) * (TM + 1) + row] = v;
                                }
                        }
                    fence_before_sync ();                       // TMEM reads ordered before the slots' next MMAs
                    __syncwarp ();
                    if (lane == 0)
                        for (int ch = 0; ch < it.nChild; ch++)
                            mbar_arrive (&barAccEmpty[(int)((u + ch) % (unsigned) NA)]);
                    u += (unsigned) it.nChild;
                    }
                if (tid == 0) TCP_T (nItem, 13);
                // the halves meet: row maxima, then the copy-out
                sMax[grp][row] = mx;
                tcp_bar_epilogue ();
                const bool doScale = it.sw >= 0;
                {
                float *dstBase = ctx.partials + (size_t)(it.dest - ctx.tipCount) * bufStride;
                constexpr int nq = SPC / 4, NCP = (nq + 1) / 2;
                // the reciprocals of the rows this thread copies, all of them first (straight-line code below: the
                // shared-memory loads of the copy are then in flight together)
                float fr[NCP];
                #pragma unroll
                for (int n = 0; n < NCP; n++)
                    {
                    const int idx = n * 256 + tid;
                    const int r = (idx < nq * TM) ? idx / nq : 0;
                    fr[n] = doScale ? __frcp_rn (fmaxf (sMax[0][r], sMax[1][r])) : 1.0f;      // = 1.0f / max, IEEE
                    }
                for (int k = 0; k < K; k++)
                    {
                    float4 *dst = reinterpret_cast<float4 *>(dstBase + ((size_t)k * C + c0) * SPC);
                    float4 v[NCP];
                    #pragma unroll
                    for (int n = 0; n < NCP; n++)
                        {
                        const int idx = n * 256 + tid;
                        const int r = (idx < nq * TM) ? idx / nq : 0, q = idx % nq;
                        v[n] = sStage[((size_t)k * NQ + ((q < NQ) ? q : 0)) * (TM + 1) + r];
                        if (q >= NQ) v[n] = make_float4 (0.f, 0.f, 0.f, 0.f);
                        }
                    #pragma unroll
                    for (int n = 0; n < NCP; n++)
                        {
                        const int idx = n * 256 + tid;
                        const int r = idx / nq;
                        v[n].x *= fr[n]; v[n].y *= fr[n]; v[n].z *= fr[n]; v[n].w *= fr[n];
                        if (r < np && idx < nq * TM)
                            dst[idx] = v[n];
                        }
                    }
                }
                if (doScale && grp == 0 && active)              // node scaler (CondLikeScaler_Gen_SSE: log in double, cast to float, src/likelihood.c:5055)
                    ctx.scalers[(size_t)it.sw * C + c] = log_of_max (fmaxf (sMax[0][row], sMax[1][row]));
                // publish: the barrier orders every epilogue thread's stores before thread 0's hand-over; the publisher warp
                // acquires it and release-stores the flag (cumulative at GPU scope), so no epilogue warp sits in a memory
                // barrier while the next item's accumulators are waiting
                if (tid == 0) TCP_T (nItem, 14);
                tcp_bar_epilogue ();
                if (tid == 0)
                    {
                    tcp_wait (&barPubEmpty[pslot], pph ^ 1);
                    sPub[pslot] = flagRow + it.oi;
                    mbar_arrive (&barPubFull[pslot]);
                    if (++pslot == TCP_NPUB) { pslot = 0; pph ^= 1; }
                    TCP_T (nItem, 15);
                    }
                continue;
                }

            // ---------------------------------------------------------------- closing item of (e, t): warps 0-3
            if (grp != 0)
                continue;
            const int nOp = ev->nOp;
            float site = 0.0f;
            if (active)
                {
                site = (ev->siteSrc >= 0) ? ctx.scalers[(size_t)ev->siteSrc * C + c] : 0.0f;
                for (int o = 0; o < nOp; o++)                  // the caller's operation order (src/likelihood.c:7938-7965)
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
                const double *rates = dvals + ev->dOff;
                const double *catW = rates + K, *freqs = rates + 2*K;
                const float *rootBase = ctx.partials + (size_t)(ev->root - ctx.tipCount) * bufStride;
                double like = 0.0;
                for (int k = 0; k < K; k++)
                    {
                    const float4 *rp = reinterpret_cast<const float4 *>(rootBase + ((size_t)k * C + c) * Sp);
                    double s = 0.0;
                    #pragma unroll 4
                    for (int q = 0; q < NQ; q++)
                        {
                        const float4 v = __ldcg (rp + q);
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
            if (lane == 0) { qSum[grp][warp & 3] = term; qAb[grp][warp & 3] = abortFlag; }
            tcp_bar_group (grp);
            if (gtid == 0)
                {
                const double s = qSum[grp][0] + qSum[grp][1] + qSum[grp][2] + qSum[grp][3];
                const int    a = qAb[grp][0] | qAb[grp][1] | qAb[grp][2] | qAb[grp][3];
                ctx.tilePartial[(size_t)it.e * numTiles + it.t] = s;
                ctx.tileAbort  [(size_t)it.e * numTiles + it.t] = a;
                __threadfence ();
                const unsigned int tk = atomicAdd (&ctx.ticket[it.e], 1u);
                if (tk == (unsigned int) numTiles - 1u)
                    {
                    __threadfence ();
                    double tot = 0.0; int ab = 0;
                    for (int tIdx = 0; tIdx < numTiles; tIdx++)
                        {
                        tot += __ldcg (&ctx.tilePartial[(size_t)it.e * numTiles + tIdx]);
                        ab  |= __ldcg (&ctx.tileAbort  [(size_t)it.e * numTiles + tIdx]);
                        }
                    if (*((volatile int *) Q.error)) ab = 1;
                    const double lnL = ab ? -DBL_MAX : tot;
                    int4 pkt;
                    pkt.x = __double2loint (lnL); pkt.y = __double2hiint (lnL); pkt.z = ab ? 1 : 0; pkt.w = seq;
                    *reinterpret_cast<int4 *>(&out[it.e]) = pkt;
                    ctx.ticket[it.e] = 0u;
                    }
                }
            tcp_bar_group (grp);                               // qSum / qAb free for the group's next closing item
            }
        }

    fence_before_sync ();
    __syncthreads ();
    if (warp == 0)
        tmem_dealloc<512> (tmemBase);
}
