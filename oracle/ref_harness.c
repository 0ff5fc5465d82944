/*
 * ref_harness.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * Linked with the UNMODIFIED reference objects (compiled from /root/reference/src by
 * oracle/Makefile) using GNU ld's --wrap on the single cross-TU call site of the
 * hot path, LogLike -> LaunchLogLikeForDivision (src/mcmc.c:7437 ->
 * src/likelihood.c:7851), and on CalcCijk (src/utils.c:9734) to see V and V^-1.
 *
 * MB200_MODE selects what the wrapper does for every likelihood evaluation:
 *   cpu     reference path only; time it and count CL updates (CPU baseline)
 *   dump    reference path drives the chain; the seam's engine calls are RECORDED
 *           (not executed) together with the reference lnL -> golden vectors
 *   shadow  reference path drives the chain; the same calls also run on the GPU
 *           engine and |lnL_gpu - lnL_cpu| / |lnL_cpu| is checked per evaluation
 *   gpu     the GPU engine alone drives the chain (the drop-in), timed
 *   oracle  like gpu, but the seam talks to the CPU oracle (oracle/liboracle.so, loaded at run time) instead of
 *           the engine: the whole host side -- seam, flip protocol, chain batching -- runs where there is no GPU,
 *           and must reproduce the reference's own trajectory (the oracle is bit-exact on the FMA build)
 * Other environment variables:
 *   MB200_DUMP_FILE   output of dump mode (default mb200_golden.bin)
 *   MB200_DUMP_MAX    stop recording after this many evaluations (default: all)
 *   MB200_TOL         shadow-mode relative tolerance (default 1e-6)
 *   MB200_REPORT      file the JSON summary is appended to (default: stderr)
 *   MB200_BATCH       chain-batched generations (binaries linked with the patched RunChain, oracle/patch_runchain.py:
 *                     mb_b200_batched): 1 (default in gpu / oracle mode) = all local chains of a generation in one
 *                     engine call; 0 = evaluate at queue time (reproduces the serial binary bit for bit)
 *   MB200_ORACLE_LIB  oracle mode: path of liboracle.so (default: next to this binary's parent directory)
 *   MB200_VIA         "fnptr": every mode reaches the engine through the node-granular function-pointer
 *                     forms installed in ModelInfo (what SetLikeFunctions would do), with the reference's
 *                     own LaunchLogLikeForDivision loop driving them; default: the seam's own loop
 *
 * Golden file: "MB200GLD" u32 version=1, then chunks {u32 tag, u32 bytes, payload}:
 *   'INST' i32 division, 12 x i32 mb200_instance_config
 *   'TIPS' i32 division, i32 tip, i32 C, C x u64
 *   'WGHT' i32 division, i32 row, i32 C, C x f32
 *   'PSTA' i32 division, i32 C, i32 matrix_length, i32 dummy_patterns, i32 uncompressed_sites, i32 freq_length,
 *          C x i32 nStates, C x i32 tiIndex, C x i32 bsIndex     (variable-state divisions; EVAL's S field
 *          is then freq_length, the number of state frequencies recorded)
 *   'EIGN' i32 division, i32 eigen, i32 S, f64 lambda[S], V[S*S], Vinv[S*S]
 *   'CIJK' i32 division, i32 eigen, i32 S, f64 block[2S+S^3]   (when V/Vinv unavailable; eigen = -2:
 *          the block travels inline with the NEXT 'EVAL' record)
 *   'EVAL' i32 division, i32 chain, i32 nMat, i32 nOp, i32 siteDst, i32 siteSrc,
 *          i32 root, i32 weightsRow, i32 flags, i32 hasPInvar, i32 K, i32 S,
 *          f64 pInvar, f64 rates[K], f64 weights[K], f64 freqs[S],
 *          nMat x {i32 matrix, i32 eigen, f64 length}, nOp x 9 x i32,
 *          f64 lnL_reference, i32 abort, i32 pad
 */
#include "bayes.h"
#include "likelihood.h"
#include "mcmc.h"
#include "model.h"
#include "utils.h"
#include "mb200.h"
#include "mb200_seam.h"

#include <time.h>
#include <dlfcn.h>
#include <unistd.h>

void __real_LaunchLogLikeForDivision (int chain, int d, MrBFlt *lnL);
MrBFlt __real_LogLike (int chain);
void __real_CalcCijk (int dim, MrBFlt *c_ijk, MrBFlt **u, MrBFlt **v);

enum { MODE_CPU, MODE_DUMP, MODE_SHADOW, MODE_GPU, MODE_ORACLE };
#define ENGINE_DRIVES(mode) ((mode) == MODE_GPU || (mode) == MODE_ORACLE)

static int        hMode = -1;
static int        hMultiPart = 1;    /* gpu mode: all divisions of a chain in flight together (MB200LogLike) */
static int        hViaFn = 0;        /* MB200_VIA=fnptr: reach the engine through the node-granular function pointers
                                        (TiProbs_B200 ... Likelihood_B200 installed in ModelInfo) driven by the
                                        reference's own LaunchLogLikeForDivision loop, instead of the seam's own loop */
extern int MB200RC_patched __attribute__((weak));   /* defined by the patched RunChain only */
static int        hBatch = 1;        /* MB200_BATCH */
static int        hBatchActive = NO; /* this generation runs chain-batched */
static long long  hFlushes = 0, hBatchedGens = 0;
static double     hSecQueue = 0.0, hSecFlush = 0.0, hSecFinish = 0.0;   /* chain-batched generations: where sec_gpu goes */
static double     hSecInit = 0.0;    /* calls that created an engine instance (CUDA context, buffers, tip upload): not in sec_gpu */
static FILE      *hDump = NULL;
static long       hDumpMax = -1, hDumped = 0;
static double     hTol = 1e-6;
static double     hSecCpu = 0.0, hSecGpu = 0.0;
static long long  hCalls = 0, hUpdates = 0, hNodeUpdates = 0, hAborts = 0, hUnsupported = 0;
static double     hMaxRel = 0.0, hSumRel = 0.0;
static long long  hCompared = 0, hFailed = 0;
static unsigned long long hLnlHash = 1469598103934665603ULL;   /* FNV-1a over every lnL handed back to the chain */

static void HashLnl (double v)
{
    const unsigned char *b = (const unsigned char *) &v;
    int i;
    for (i=0; i<8; i++)
        hLnlHash = (hLnlHash ^ b[i]) * 1099511628211ULL;
}

/* last eigensystem seen by CalcCijk */
static int        hEigDim = 0;
static double    *hEigU = NULL, *hEigV = NULL;
#define HRING 8                       /* the last few CalcCijk inputs: multi-part slots (NY98) call it once per part */
static double    *hRingU[HRING], *hRingV[HRING];
static int        hRingDim[HRING], hRingPos = 0;

/* evaluation stashed by the recorder until the reference value is known */
static struct
    {
    int                  valid, instance;
    mb200_evaluation     ev;
    mb200_operation     *ops;
    mb200_matrix_update *mats;
    int                  capOps, capMats;
    double               lnLGpu;
    int                  statusGpu;
    double              *inlineEig;
    } hLast;

static int hInstDivision[4096];     /* recorder instance id -> division */
static int hInstReal[4096];         /* recorder instance id -> engine instance (shadow) */
static mb200_instance_config hInstCfg[4096];
static int hInstFreqLen[4096];     /* variable-state instances: entries of state_freqs in use */
static int hNumInst = 0;

static double Now (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void Chunk (const char *tag, const void *hdr, size_t hdrBytes, const void *body, size_t bodyBytes)
{
    unsigned int n = (unsigned int)(hdrBytes + bodyBytes);
    if (!hDump)
        return;
    fwrite (tag, 1, 4, hDump);
    fwrite (&n, 4, 1, hDump);
    if (hdrBytes)  fwrite (hdr, 1, hdrBytes, hDump);
    if (bodyBytes) fwrite (body, 1, bodyBytes, hDump);
}

/* ---- what shadow mode runs beside the reference: the engine, or (MB200_SHADOW_BACKEND=oracle) the CPU oracle -- the
 *      latter lets the seam's policies (dynamic rescaling, its retry) be checked per evaluation where there is no GPU */
static struct
    {
    int (*create) (const mb200_instance_config *, int *);
    int (*finalize) (int);
    int (*tips) (int, int, const uint64_t *);
    int (*weights) (int, int, const float *);
    int (*cijk) (int, int, const double *);
    int (*eval) (int, const mb200_evaluation *, int, double *, int *);
    int (*pstates) (int, const int *, const int *, const int *, int, int, int);
    int (*setm) (int, int, const float *);
    } hSh = { mb200_create_instance, mb200_finalize_instance, mb200_set_tip_states, mb200_set_pattern_weights, mb200_set_cijk,
              mb200_evaluate, mb200_set_pattern_states, mb200_set_transition_matrix };

/* ---- recording backend ------------------------------------------------------------- */
static int rec_create (const mb200_instance_config *c, int *inst)
{
    int hdr[13], id = hNumInst++;
    /* division is found by matching the seam's bookkeeping after the call; the seam
       creates instances one division at a time, in LogLike order */
    hInstCfg[id] = *c;
    hInstReal[id] = -1;
    *inst = id;
    if (hMode == MODE_SHADOW)
        {
        int rc = hSh.create (c, &hInstReal[id]);
        if (rc != MB200_SUCCESS)
            return rc;
        }
    hdr[0] = -1;   /* patched by the caller through hInstDivision once known */
    memcpy (hdr + 1, c, 12 * sizeof(int));
    (void) hdr;
    return MB200_SUCCESS;
}

static int rec_finalize (int inst)
{
    if (hMode == MODE_SHADOW && hInstReal[inst] >= 0)
        return hSh.finalize (hInstReal[inst]);
    return MB200_SUCCESS;
}

static int rec_tips (int inst, int tip, const uint64_t *masks)
{
    int hdr[3];
    hdr[0] = hInstDivision[inst]; hdr[1] = tip; hdr[2] = hInstCfg[inst].pattern_count;
    Chunk ("TIPS", hdr, sizeof(hdr), masks, (size_t)hdr[2] * sizeof(uint64_t));
    if (hMode == MODE_SHADOW)
        return hSh.tips (hInstReal[inst], tip, masks);
    return MB200_SUCCESS;
}

static int rec_weights (int inst, int row, const float *w)
{
    int hdr[3];
    hdr[0] = hInstDivision[inst]; hdr[1] = row; hdr[2] = hInstCfg[inst].pattern_count;
    Chunk ("WGHT", hdr, sizeof(hdr), w, (size_t)hdr[2] * sizeof(float));
    if (hMode == MODE_SHADOW)
        return hSh.weights (hInstReal[inst], row, w);
    return MB200_SUCCESS;
}

static int rec_pstates (int inst, const int *ns, const int *ti, const int *bs, int matLen, int dummy, int uncompressed)
{
    int hdr[6], c, nPat = hInstCfg[inst].pattern_count, freqLen = 0, *body;

    for (c=0; c<nPat; c++)
        if (bs[c] + ns[c] > freqLen)
            freqLen = bs[c] + ns[c];
    hInstFreqLen[inst] = freqLen;
    hdr[0] = hInstDivision[inst]; hdr[1] = nPat; hdr[2] = matLen; hdr[3] = dummy; hdr[4] = uncompressed; hdr[5] = freqLen;
    body = (int *) malloc ((size_t)3 * nPat * sizeof(int));
    memcpy (body, ns, (size_t)nPat * sizeof(int));
    memcpy (body + nPat, ti, (size_t)nPat * sizeof(int));
    memcpy (body + 2*nPat, bs, (size_t)nPat * sizeof(int));
    Chunk ("PSTA", hdr, sizeof(hdr), body, (size_t)3 * nPat * sizeof(int));
    free (body);
    if (hMode == MODE_SHADOW)
        return hSh.pstates (hInstReal[inst], ns, ti, bs, matLen, dummy, uncompressed);
    return MB200_SUCCESS;
}

static int rec_cijk (int inst, int eigen, const double *block)
{
    int     hdr[3], S = hInstCfg[inst].state_count, i, j, k, same = 0;
    size_t  n3 = (size_t)S * S * S;

    const int parts = (((hInstCfg[inst].flags >> 8) & 0xff) > 1) ? ((hInstCfg[inst].flags >> 8) & 0xff) : 1;

    hdr[0] = hInstDivision[inst]; hdr[1] = eigen; hdr[2] = S;
    if (hDump)
        {
        /* prefer the compact (lambda, V, V^-1) form when it reproduces the block exactly */
        if (parts == 1 && hEigDim == S && hEigU != NULL)
            {
            const double *c = block + 2*S;
            same = 1;
            for (i=0; i<S && same; i++)
                for (j=0; j<S && same; j++)
                    for (k=0; k<S; k++)
                        if (c[((size_t)i*S + j)*S + k] != hEigU[i*S+k] * hEigV[k*S+j])
                            { same = 0; break; }
            }
        if (same)
            {
            double *buf = (double *) malloc ((size_t)(S + 2*S*S) * sizeof(double));
            memcpy (buf, block, (size_t)S * sizeof(double));
            memcpy (buf + S, hEigU, (size_t)S*S * sizeof(double));
            memcpy (buf + S + S*S, hEigV, (size_t)S*S * sizeof(double));
            Chunk ("EIGN", hdr, sizeof(hdr), buf, (size_t)(S + 2*S*S) * sizeof(double));
            free (buf);
            }
        else if (parts > 1)
            {
            /* one (lambda, V, V^-1) triple per part when the recent CalcCijk inputs reproduce every part
               exactly (they do: UpDateCijk calls CalcCijk once per omega category); else the raw blocks */
            int     found[32], p, r, ok = 1;
            const size_t partLen = 2*(size_t)S + n3;
            for (p=0; p<parts && p<32 && ok; p++)
                {
                const double *c = block + (size_t)p * partLen + 2*S;
                found[p] = -1;
                for (r=0; r<HRING && found[p] < 0; r++)
                    {
                    int match = (hRingDim[r] == S && hRingU[r] != NULL);
                    for (i=0; i<S && match; i++)
                        for (j=0; j<S && match; j++)
                            for (k=0; k<S; k++)
                                if (c[((size_t)i*S + j)*S + k] != hRingU[r][i*S+k] * hRingV[r][k*S+j])
                                    { match = 0; break; }
                    if (match) found[p] = r;
                    }
                if (found[p] < 0) ok = 0;
                }
            if (ok && parts <= 32)
                for (p=0; p<parts; p++)
                    {
                    int ph[3] = { hdr[0], eigen | (p << 16) | (parts << 24), S };
                    double *buf = (double *) malloc ((size_t)(S + 2*S*S) * sizeof(double));
                    memcpy (buf, block + (size_t)p * partLen, (size_t)S * sizeof(double));
                    memcpy (buf + S, hRingU[found[p]], (size_t)S*S * sizeof(double));
                    memcpy (buf + S + S*S, hRingV[found[p]], (size_t)S*S * sizeof(double));
                    Chunk ("EIGN", ph, sizeof(ph), buf, (size_t)(S + 2*S*S) * sizeof(double));
                    free (buf);
                    }
            else
                Chunk ("CIJK", hdr, sizeof(hdr), block, (size_t)parts * partLen * sizeof(double));
            }
        else
            Chunk ("CIJK", hdr, sizeof(hdr), block, (2*(size_t)S + n3) * sizeof(double));
        }
    if (hMode == MODE_SHADOW)
        return hSh.cijk (hInstReal[inst], eigen, block);
    return MB200_SUCCESS;
}

/* host-built transition matrices (STANDARD divisions with ordered characters / unequal frequencies): shadow runs pass them
   on; the golden record format has no such record, so dump runs leave those divisions to the reference */
static int rec_setm (int inst, int matrix, const float *in)
{
    return hSh.setm (inst, matrix, in);
}

static int rec_noread (int inst, int index, float *out)
{
    (void) inst; (void) index; (void) out;
    return MB200_ERROR_UNSUPPORTED;
}

static int rec_eval (int inst, const mb200_evaluation *e, int n, double *lnL, int *status)
{
    int i;
    if (n != 1)
        return MB200_ERROR_UNSUPPORTED;
    if (e->operation_count > hLast.capOps)
        {
        hLast.capOps = e->operation_count + 64;
        hLast.ops = (mb200_operation *) realloc (hLast.ops, (size_t)hLast.capOps * sizeof(mb200_operation));
        }
    if (e->matrix_update_count > hLast.capMats)
        {
        hLast.capMats = e->matrix_update_count + 64;
        hLast.mats = (mb200_matrix_update *) realloc (hLast.mats, (size_t)hLast.capMats * sizeof(mb200_matrix_update));
        }
    hLast.ev = *e;
    for (i=0; i<e->operation_count; i++)      hLast.ops[i]  = e->operations[i];
    for (i=0; i<e->matrix_update_count; i++)  hLast.mats[i] = e->matrix_updates[i];
    hLast.ev.operations = hLast.ops;
    hLast.ev.matrix_updates = hLast.mats;
    if (e->inline_eigen != NULL)
        {
        int S = hInstCfg[inst].state_count;
        size_t nb = (2*(size_t)S + (size_t)S*S*S) * sizeof(double);
        hLast.inlineEig = (double *) realloc (hLast.inlineEig, nb);
        memcpy (hLast.inlineEig, e->inline_eigen, nb);
        hLast.ev.inline_eigen = hLast.inlineEig;
        }
    hLast.instance = inst;
    hLast.valid = 1;
    lnL[0] = 0.0;
    status[0] = MB200_EVAL_OK;
    if (hMode == MODE_SHADOW)
        {
        double t0 = Now ();
        int rc = hSh.eval (hInstReal[inst], e, 1, lnL, status);
        hSecGpu += Now () - t0;
        hLast.lnLGpu = lnL[0];
        hLast.statusGpu = status[0];
        return rc;
        }
    return MB200_SUCCESS;
}

static void WriteEval (int division, int chain, double lnLRef, int aborted)
{
    int         hdr[12], tail[2], i;
    size_t      nb;
    char       *buf, *q;
    const mb200_evaluation *e = &hLast.ev;
    int         K = hInstCfg[hLast.instance].category_count, S = hInstCfg[hLast.instance].state_count;

    if (!hDump || !hLast.valid)
        return;
    if (e->inline_eigen != NULL)
        {
        int eh[3];
        eh[0] = division; eh[1] = MB200_EIGEN_INLINE; eh[2] = S;
        Chunk ("CIJK", eh, sizeof(eh), e->inline_eigen, (2*(size_t)S + (size_t)S*S*S) * sizeof(double));
        }
    hdr[0] = division; hdr[1] = chain; hdr[2] = e->matrix_update_count; hdr[3] = e->operation_count;
    hdr[4] = e->site_scaler_dst; hdr[5] = e->site_scaler_src; hdr[6] = e->root_buffer; hdr[7] = e->weights_row;
    if (hInstCfg[hLast.instance].flags & MB200_CONFIG_VARIABLE_STATES)
        S = hInstFreqLen[hLast.instance];           /* how many state frequencies the division uses */
    hdr[8] = e->flags; hdr[9] = e->has_p_invar; hdr[10] = K; hdr[11] = S;
    nb = sizeof(double) * (size_t)(1 + 2*K + S) + (size_t)e->matrix_update_count * 16 + (size_t)e->operation_count * 36 + 16;
    buf = q = (char *) malloc (nb);
    memcpy (q, &e->p_invar, 8); q += 8;
    memcpy (q, e->category_rates, 8*(size_t)K); q += 8*(size_t)K;
    memcpy (q, e->category_weights, 8*(size_t)K); q += 8*(size_t)K;
    memcpy (q, e->state_freqs, 8*(size_t)S); q += 8*(size_t)S;
    for (i=0; i<e->matrix_update_count; i++)
        {
        memcpy (q, &e->matrix_updates[i].matrix, 4); q += 4;
        memcpy (q, &e->matrix_updates[i].eigen, 4); q += 4;
        memcpy (q, &e->matrix_updates[i].length, 8); q += 8;
        }
    for (i=0; i<e->operation_count; i++)
        {
        memcpy (q, &e->operations[i], 36); q += 36;
        }
    memcpy (q, &lnLRef, 8); q += 8;
    tail[0] = aborted; tail[1] = 0;
    memcpy (q, tail, 8); q += 8;
    Chunk ("EVAL", hdr, sizeof(hdr), buf, (size_t)(q - buf));
    free (buf);
    hDumped++;
}

/* ---- summary ----------------------------------------------------------------------- */
static void Report (void)
{
    const char *path = getenv ("MB200_REPORT");
    const char *names[] = { "cpu", "dump", "shadow", "gpu", "oracle" };
    FILE *f = path ? fopen (path, "a") : stderr;
    double cjH = 0.0, cjU = 0.0;
    long long cjN = 0;
    if (!f) f = stderr;
    MB200SeamCijkTimes (&cjH, &cjU, &cjN);
    fprintf (f, "{\"mb200_harness\": \"%s\", \"calls\": %lld, \"node_updates\": %lld, \"cl_updates\": %lld, "
                "\"sec_cpu\": %.6f, \"sec_gpu\": %.6f, \"aborts\": %lld, \"unsupported_calls\": %lld, "
                "\"compared\": %lld, \"failed\": %lld, \"max_rel\": %.3e, \"mean_rel\": %.3e, \"tol\": %.1e, \"dumped\": %ld, "
                "\"via\": \"%s\", \"lnl_hash\": \"%016llx\", \"batched_generations\": %lld, \"flushes\": %lld, \"rescale_retries\": %lld, \"sec_queue\": %.6f, \"sec_flush\": %.6f, \"sec_finish\": %.6f, \"sec_init\": %.6f, \"sec_cijk_host\": %.6f, \"sec_cijk_upload\": %.6f, \"cijk_updates\": %lld, \"device_eigens\": %lld}\n",
             names[hMode], hCalls, hNodeUpdates, hUpdates, hSecCpu, hSecGpu, hAborts, hUnsupported,
             hCompared, hFailed, (hMaxRel == hMaxRel && hMaxRel < 1e300) ? hMaxRel : 9.999e99,
             (hCompared && hSumRel == hSumRel && hSumRel < 1e300) ? hSumRel / hCompared : (hCompared ? 9.999e99 : 0.0), hTol, hDumped,
             hViaFn ? "fnptr" : "seam", hLnlHash, hBatchedGens, hFlushes, MB200SeamRescaleRetries (), hSecQueue, hSecFlush, hSecFinish, hSecInit, cjH, cjU, cjN, MB200SeamDeviceEigens ());
    if (f != stderr) fclose (f);
    if (hDump) { fclose (hDump); hDump = NULL; }
    if (hMode == MODE_SHADOW || ENGINE_DRIVES (hMode))
        MB200SeamFinalize ();
}

/* ---- oracle backend: the seam's calls go to the CPU restatement (same C-ABI, orc_ prefix) --------- */
static struct
    {
    void *lib;
    int (*create) (const mb200_instance_config *, int *);
    int (*finalize) (int);
    int (*arith) (int, int);
    int (*tips) (int, int, const uint64_t *);
    int (*weights) (int, int, const float *);
    int (*cijk) (int, int, const double *);
    int (*eval) (int, const mb200_evaluation *, int, double *, int *);
    int (*pstates) (int, const int *, const int *, const int *, int, int, int);
    int (*getp) (int, int, float *);
    int (*getm) (int, int, float *);
    int (*gets) (int, int, float *);
    int (*setm) (int, int, const float *);
    } hOrc;

static int orc_create (const mb200_instance_config *c, int *inst)
{
    int rc = hOrc.create (c, inst);
    int fma = 0;
#   if defined (HAVE_FMA3) || defined (FMA_ENABLED)
    fma = 1;                                            /* ORC_ARITH_FMA: this binary's reference objects use the FMA kernels ... */
#   endif
    if (c->flags & MB200_CONFIG_SCALAR_KERNELS)
        fma = 0;                                        /* ... except where the reference itself falls back to its scalar kernels */
    if (rc == MB200_SUCCESS) hOrc.arith (*inst, fma);
    return rc;
}

static void LoadOracle (void)
{
    const char *path = getenv ("MB200_ORACLE_LIB");
    char        buf[4096];

    if (!path)
        {
        ssize_t n = readlink ("/proc/self/exe", buf, sizeof(buf) - 64);
        char   *slash;
        if (n <= 0) { fprintf (stderr, "oracle mode: cannot locate the binary\n"); exit (2); }
        buf[n] = 0;
        slash = strrchr (buf, '/');  if (slash) *slash = 0;      /* .../oracle/_ref */
        slash = strrchr (buf, '/');  if (slash) *slash = 0;      /* .../oracle */
        strcat (buf, "/liboracle.so");
        path = buf;
        }
    hOrc.lib = dlopen (path, RTLD_NOW);
    if (!hOrc.lib) { fprintf (stderr, "oracle mode: %s\n", dlerror ()); exit (2); }
    hOrc.create   = (int (*) (const mb200_instance_config *, int *)) dlsym (hOrc.lib, "orc_create_instance");
    hOrc.finalize = (int (*) (int)) dlsym (hOrc.lib, "orc_finalize_instance");
    hOrc.arith    = (int (*) (int, int)) dlsym (hOrc.lib, "orc_set_arith");
    hOrc.tips     = (int (*) (int, int, const uint64_t *)) dlsym (hOrc.lib, "orc_set_tip_states");
    hOrc.weights  = (int (*) (int, int, const float *)) dlsym (hOrc.lib, "orc_set_pattern_weights");
    hOrc.cijk     = (int (*) (int, int, const double *)) dlsym (hOrc.lib, "orc_set_cijk");
    hOrc.eval     = (int (*) (int, const mb200_evaluation *, int, double *, int *)) dlsym (hOrc.lib, "orc_evaluate");
    hOrc.pstates  = (int (*) (int, const int *, const int *, const int *, int, int, int)) dlsym (hOrc.lib, "orc_set_pattern_states");
    hOrc.getp     = (int (*) (int, int, float *)) dlsym (hOrc.lib, "orc_get_partials");
    hOrc.getm     = (int (*) (int, int, float *)) dlsym (hOrc.lib, "orc_get_transition_matrix");
    hOrc.gets     = (int (*) (int, int, float *)) dlsym (hOrc.lib, "orc_get_scalers");
    hOrc.setm     = (int (*) (int, int, const float *)) dlsym (hOrc.lib, "orc_set_transition_matrix");
    if (!hOrc.create || !hOrc.finalize || !hOrc.arith || !hOrc.tips || !hOrc.weights || !hOrc.cijk || !hOrc.eval || !hOrc.pstates)
        { fprintf (stderr, "oracle mode: %s lacks part of the orc_ API\n", path); exit (2); }
}

static void Setup (void)
{
    const char *s = getenv ("MB200_MODE");
    hMode = MODE_CPU;
    if (s && !strcmp (s, "dump"))   hMode = MODE_DUMP;
    if (s && !strcmp (s, "shadow")) hMode = MODE_SHADOW;
    if (s && !strcmp (s, "gpu"))    hMode = MODE_GPU;
    if (s && !strcmp (s, "oracle")) hMode = MODE_ORACLE;
    if ((s = getenv ("MB200_BATCH")) != NULL) hBatch = atoi (s);
    if ((s = getenv ("MB200_TOL")) != NULL)      hTol = atof (s);
    if ((s = getenv ("MB200_MULTIPART")) != NULL) hMultiPart = atoi (s);
    if ((s = getenv ("MB200_VIA")) != NULL && !strcmp (s, "fnptr")) { hViaFn = 1; hMultiPart = 0; }
    if ((s = getenv ("MB200_DUMP_MAX")) != NULL) hDumpMax = atol (s);
    memset (&hLast, 0, sizeof(hLast));
    if (hMode == MODE_DUMP)
        {
        unsigned int ver = 1;
        s = getenv ("MB200_DUMP_FILE");
        hDump = fopen (s ? s : "mb200_golden.bin", "wb");
        if (!hDump) { perror ("MB200_DUMP_FILE"); exit (2); }
        fwrite ("MB200GLD", 1, 8, hDump);
        fwrite (&ver, 4, 1, hDump);
        }
    if (hMode == MODE_SHADOW && (s = getenv ("MB200_SHADOW_BACKEND")) != NULL && !strcmp (s, "oracle"))
        {
        LoadOracle ();
        hSh.create = orc_create;   hSh.finalize = hOrc.finalize; hSh.tips = hOrc.tips; hSh.weights = hOrc.weights;
        hSh.cijk = hOrc.cijk;      hSh.eval = hOrc.eval;         hSh.pstates = hOrc.pstates;
        hSh.setm = hOrc.setm;
        }
    if (hMode == MODE_DUMP || hMode == MODE_SHADOW)
        {
        MB200SeamBackend be = { rec_create, rec_finalize, rec_tips, rec_weights, rec_cijk, rec_eval, NULL, NULL, rec_pstates, NULL, NULL, NULL };
        if (hMode == MODE_SHADOW)
            {
            /* the reference drives and its own readers see its own host arrays (the wrappers are not installed): a nominal
               read-back lets divisions that report ancestral states / site rates be shadowed like any other */
            be.get_partials = rec_noread; be.get_transition_matrix = rec_noread; be.get_scalers = rec_noread;
            be.set_transition_matrix = rec_setm;
            }
        MB200SeamSetBackend (&be);
        }
    if (hMode == MODE_ORACLE)
        {
        MB200SeamBackend be;
        memset (&be, 0, sizeof(be));       /* optional entry points the oracle has no counterpart of stay NULL */
        LoadOracle ();
        be.create_instance = orc_create;       be.finalize_instance = hOrc.finalize;
        be.set_tip_states = hOrc.tips;         be.set_pattern_weights = hOrc.weights;
        be.set_cijk = hOrc.cijk;               be.evaluate = hOrc.eval;
        be.evaluate_begin = NULL;              be.evaluate_end = NULL;
        be.set_pattern_states = hOrc.pstates;
        be.get_partials = hOrc.getp; be.get_transition_matrix = hOrc.getm; be.get_scalers = hOrc.gets;
        be.set_transition_matrix = hOrc.setm;
        MB200SeamSetBackend (&be);
        }
    if (ENGINE_DRIVES (hMode) && hBatch && &MB200RC_patched != NULL)
        MB200BatchEnable (YES);                 /* before the first instance is created */
    atexit (Report);
}

void __wrap_CalcCijk (int dim, MrBFlt *c_ijk, MrBFlt **u, MrBFlt **v)
{
    int i, j;
    if (hMode == MODE_DUMP)
        {
        if (dim != hEigDim)
            {
            free (hEigU); free (hEigV);
            hEigU = (double *) malloc ((size_t)dim*dim*sizeof(double));
            hEigV = (double *) malloc ((size_t)dim*dim*sizeof(double));
            hEigDim = dim;
            }
        for (i=0; i<dim; i++)
            for (j=0; j<dim; j++)
                {
                hEigU[i*dim+j] = u[i][j];
                hEigV[i*dim+j] = v[i][j];
                }
        {
        const int r = hRingPos % HRING;
        if (hRingDim[r] != dim)
            {
            free (hRingU[r]); free (hRingV[r]);
            hRingU[r] = (double *) malloc ((size_t)dim*dim*sizeof(double));
            hRingV[r] = (double *) malloc ((size_t)dim*dim*sizeof(double));
            hRingDim[r] = dim;
            }
        memcpy (hRingU[r], hEigU, (size_t)dim*dim*sizeof(double));
        memcpy (hRingV[r], hEigV, (size_t)dim*dim*sizeof(double));
        hRingPos++;
        }
        }
    __real_CalcCijk (dim, c_ijk, u, v);
}

/* ---- index-table snapshot (so the reference can redo the same flips) ---------------- */
typedef struct
    {
    int  nNodes, *cl, *clS, *ti, *tiS, *ns, *nsS, *un, *unS, site, siteS;
    } Snap;

static void SnapTake (Snap *s, ModelInfo *m, int chain, int nNodes)
{
    size_t nb = (size_t)nNodes * sizeof(int);
    s->nNodes = nNodes;
    s->cl  = (int *) malloc (8*nb);
    s->clS = s->cl + nNodes;  s->ti  = s->clS + nNodes; s->tiS = s->ti + nNodes;
    s->ns  = s->tiS + nNodes; s->nsS = s->ns + nNodes;  s->un  = s->nsS + nNodes; s->unS = s->un + nNodes;
    memcpy (s->cl,  m->condLikeIndex[chain], nb);    memcpy (s->clS, m->condLikeScratchIndex, nb);
    memcpy (s->ti,  m->tiProbsIndex[chain], nb);     memcpy (s->tiS, m->tiProbsScratchIndex, nb);
    memcpy (s->ns,  m->nodeScalerIndex[chain], nb);  memcpy (s->nsS, m->nodeScalerScratchIndex, nb);
    memcpy (s->un,  m->unscaledNodes[chain], nb);    memcpy (s->unS, m->unscaledNodesScratch, nb);
    s->site = m->siteScalerIndex[chain]; s->siteS = m->siteScalerScratchIndex;
}

static void SnapRestore (Snap *s, ModelInfo *m, int chain)
{
    size_t nb = (size_t)s->nNodes * sizeof(int);
    memcpy (m->condLikeIndex[chain], s->cl, nb);     memcpy (m->condLikeScratchIndex, s->clS, nb);
    memcpy (m->tiProbsIndex[chain], s->ti, nb);      memcpy (m->tiProbsScratchIndex, s->tiS, nb);
    memcpy (m->nodeScalerIndex[chain], s->ns, nb);   memcpy (m->nodeScalerScratchIndex, s->nsS, nb);
    memcpy (m->unscaledNodes[chain], s->un, nb);     memcpy (m->unscaledNodesScratch, s->unS, nb);
    m->siteScalerIndex[chain] = s->site; m->siteScalerScratchIndex = s->siteS;
    free (s->cl);
}

static long long CountDirty (Tree *t)
{
    int i; long long n = 0;
    for (i=0; i<t->nIntNodes; i++)
        if (t->intDownPass[i]->upDateCl == YES)
            n++;
    return n;
}

/* LogLike (src/mcmc.c:7396): in gpu mode the division loop is the seam's partition-batched one */
static void CpuPathTimed (int chain, int d, MrBFlt *lnL)
{
    hUnsupported++;
    __real_LaunchLogLikeForDivision (chain, d, lnL);
}

MrBFlt __wrap_LogLike (int chain)
{
    int     d;
    double  t0;
    MrBFlt  v;

    if (hMode < 0)
        Setup ();
    if (!ENGINE_DRIVES (hMode) || hMultiPart == 0 || chainParams.runWithData == NO)
        {
        v = __real_LogLike (chain);
        HashLnl (v);
        return v;
        }
    for (d=0; d<numCurrentDivisions; d++)
        {
        ModelInfo *m = &modelSettings[d];
        if (m->upDateCl == YES)
            {
            long long dirty = CountDirty (GetTree (m->brlens, chain, state[chain]));
            hCalls++;
            hNodeUpdates += dirty;
            hUpdates += dirty * m->numChars * m->numRateCats * m->numOmegaCats;
            }
        }
    t0 = Now ();
    v = MB200LogLike (chain, CpuPathTimed);
    hSecGpu += Now () - t0;
    if (abortMove == YES) hAborts++;
    HashLnl (v);
    return v;
}

/* ---- hooks of the patched RunChain (oracle/patch_runchain.py; binaries mb_b200_batched*) ----------
 * A generation that cannot be batched (MB200RC_Begin returns NO) runs the original serial loop body. */
int MB200RC_Begin (void)
{
    if (hMode < 0)
        Setup ();
    if (ENGINE_DRIVES (hMode))
        {
        int d;
        for (d=0; d<numCurrentDivisions; d++)
            if (((modelSettings[d].printAncStates == YES || modelSettings[d].printSiteRates == YES) &&
                 modelSettings[d].PrintSiteRates != &PrintSiteRates_B200) ||
                ((modelSettings[d].printPosSel == YES || modelSettings[d].printSiteOmegas == YES) &&
                 modelSettings[d].PosSelProbs != &PosSelProbs_B200))
                MB200InstallReaders (d);
        }
    hBatchActive = (hBatch && ENGINE_DRIVES (hMode) && !hViaFn) ? MB200BatchBegin () : NO;
    if (hBatchActive == YES)
        hBatchedGens++;
    return hBatchActive;
}

void MB200RC_Enter (int chain, int phase)
{
    if (hBatchActive == YES)
        MB200BatchEnterChain (chain, phase);
}

void MB200RC_Leave (int chain, int phase)
{
    if (hBatchActive == YES)
        MB200BatchLeaveChain (chain, phase);
}

void MB200RC_Queue (int chain)
{
    int    d;
    double t0;

    if (hBatchActive == NO)
        return;
    for (d=0; d<numCurrentDivisions; d++)
        {
        ModelInfo *m = &modelSettings[d];
        if (m->upDateCl == YES)
            {
            long long dirty = CountDirty (GetTree (m->brlens, chain, state[chain]));
            hCalls++;
            hNodeUpdates += dirty;
            hUpdates += dirty * m->numChars * m->numRateCats * m->numOmegaCats;
            }
        }
    t0 = Now ();
    MB200BatchQueueLogLike (chain);
    hSecQueue += Now () - t0;
    hSecGpu += Now () - t0;
}

void MB200RC_Flush (void)
{
    double t0;

    if (hBatchActive == NO)
        return;
    t0 = Now ();
    MB200BatchFlush ();
    hSecFlush += Now () - t0;
    hSecGpu += Now () - t0;
    hFlushes++;
}

MrBFlt MB200RC_Finish (int chain)
{
    MrBFlt v;

    if (hBatchActive == NO)
        return MRBFLT_NEG_MAX;
    {
    double t0 = Now ();
    v = MB200BatchFinishLogLike (chain);
    hSecFinish += Now () - t0;
    hSecGpu += Now () - t0;
    }
    if (abortMove == YES) hAborts++;
    HashLnl (v);
    return v;
}

void __wrap_LaunchLogLikeForDivision (int chain, int d, MrBFlt *lnL)
{
    ModelInfo  *m = &modelSettings[d];
    Tree       *tree;
    long long   dirty;
    double      t0;

    if (hMode < 0)
        Setup ();

    tree  = GetTree (m->brlens, chain, state[chain]);
    dirty = CountDirty (tree);
    hCalls++;
    hNodeUpdates += dirty;
    hUpdates += dirty * m->numChars * m->numRateCats * m->numOmegaCats;

    if (hMode == MODE_CPU || (hMode == MODE_DUMP && hDumpMax >= 0 && hDumped >= hDumpMax))
        {
        t0 = Now ();
        __real_LaunchLogLikeForDivision (chain, d, lnL);
        hSecCpu += Now () - t0;
        return;
        }

    if (ENGINE_DRIVES (hMode))
        {
        const int fresh = (MB200SeamInstance (d) < 0);
        if (((m->printAncStates == YES || m->printSiteRates == YES) && m->PrintSiteRates != &PrintSiteRates_B200) ||
            ((m->printPosSel == YES || m->printSiteOmegas == YES) && m->PosSelProbs != &PosSelProbs_B200))
            MB200InstallReaders (d);            /* what SetLikeFunctions would do (it runs again for every mcmc command) */
        t0 = Now ();
        if (hViaFn)
            {
            /* SetLikeFunctions runs again for every mcmc command: (re)install when the pointer is not ours */
            if (m->Likelihood != &Likelihood_B200 && MB200InstallLikeFunctions (d) == ERROR)
                hUnsupported++;
            __real_LaunchLogLikeForDivision (chain, d, lnL);
            }
        else if (MB200LaunchLogLikeForDivision (chain, d, lnL) == NO)
            {
            hUnsupported++;
            __real_LaunchLogLikeForDivision (chain, d, lnL);
            }
        if (fresh) hSecInit += Now () - t0;
        else       hSecGpu += Now () - t0;
        if (abortMove == YES) hAborts++;
        return;
        }

    /* dump / shadow: engine calls first on a snapshot of the index tables, then the
       reference does the very same flips for real */
    {
    Snap    snap;
    int     hadCijk = m->upDateCijk, handled, nInstBefore = hNumInst, savedAbort = abortMove;
    MrBFlt  lnLSeam = 0.0, lnLRef = 0.0;

    if (MB200SeamDivisionSupported (m) == NO)
        {
        hUnsupported++;
        t0 = Now ();
        __real_LaunchLogLikeForDivision (chain, d, lnL);
        hSecCpu += Now () - t0;
        return;
        }
    if (MB200SeamInstance (d) < 0)
        {
        /* the instance about to be created belongs to this division */
        int hdr[13];
        hInstDivision[hNumInst] = d;
        hdr[0] = d;
        {
        mb200_instance_config c;
        MB200SeamDivisionConfig (m, d, &c);     /* as InitBeagleInstance in the seam ... */
        c.device = 0; c.max_evaluations = 1;    /* ... minus what depends on the machine or the run */
        memcpy (hdr + 1, &c, 12 * sizeof(int));
        }
        Chunk ("INST", hdr, sizeof(hdr), NULL, 0);
        }
    (void) nInstBefore;

    SnapTake (&snap, m, chain, tree->nNodes);
    hLast.valid = 0;
    if (hViaFn)
        {
        TiProbFxn       fT = m->TiProbs;
        LikeDownFxn     fD = m->CondLikeDown;
        LikeRootFxn     fR = m->CondLikeRoot;
        LikeScalerFxn   fS = m->CondLikeScaler;
        LikeFxn         fL = m->Likelihood;
        handled = (MB200InstallLikeFunctions (d) == NO_ERROR) ? YES : NO;
        if (handled == YES)
            __real_LaunchLogLikeForDivision (chain, d, &lnLSeam);   /* the reference's loop over OUR pointers */
        m->TiProbs = fT; m->CondLikeDown = fD; m->CondLikeRoot = fR; m->CondLikeScaler = fS; m->Likelihood = fL;
        }
    else
        handled = MB200LaunchLogLikeForDivision (chain, d, &lnLSeam);   /* runs UpDateCijk for real */
    SnapRestore (&snap, m, chain);
    abortMove = savedAbort;

    if (hadCijk == YES)
        m->upDateCijk = NO;              /* already done by the seam call above */
    t0 = Now ();
    __real_LaunchLogLikeForDivision (chain, d, &lnLRef);
    hSecCpu += Now () - t0;
    m->upDateCijk = hadCijk;
    *lnL = lnLRef;
    if (abortMove == YES) hAborts++;

    if (handled == YES && hLast.valid)
        {
        if (hMode == MODE_DUMP)
            WriteEval (d, chain, lnLRef, lnLRef == MRBFLT_NEG_MAX);
        else
            {
            double rel;
            if (lnLRef == MRBFLT_NEG_MAX || hLast.statusGpu != MB200_EVAL_OK)
                rel = (lnLRef == MRBFLT_NEG_MAX && hLast.statusGpu != MB200_EVAL_OK) ? 0.0 : 1.0;
            else
                rel = fabs (hLast.lnLGpu - lnLRef) / fabs (lnLRef);
            hCompared++;
            hSumRel += rel;
            if (rel > hMaxRel || rel != rel) hMaxRel = rel;
            if (!(rel <= hTol))
                {
                hFailed++;
                if (hFailed <= 20)
                    fprintf (stderr, "MB200 SHADOW MISMATCH call %lld chain %d div %d: gpu %.17g cpu %.17g rel %.3e\n",
                             hCalls, chain, d, hLast.lnLGpu, lnLRef, rel);
                }
            }
        }
    }
}
