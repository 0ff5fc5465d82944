/*
 * mb200_seam.c -- host-side seam between MrBayes' C code and the B200 engine.
 *
 * This translation unit is compiled TOGETHER WITH the reference's headers
 * (-I<mrbayes>/src) and linked into the reference's `mb` binary.  It
 * re-implements the accelerator entry points that the reference declares in
 * src/mbbeagle.h:13-28 (InitBeagleInstance, LaunchBEAGLELogLikeForDivision,
 * TreeTiProbs_Beagle, TreeCondLikes_Beagle_Always_Rescale,
 * TreeLikelihood_Beagle) on top of the C-ABI in include/mb200.h instead of
 * libhmsbeagle, and it follows the arithmetic rules of the BUILT-IN path
 * (src/likelihood.c:7851-7973), which is the parity target:
 *
 *   - three-neighbour update at the interior root of unrooted trees
 *     (CondLikeRoot_*, src/likelihood.c:7920-7931) instead of BEAGLE's edge
 *     likelihood (src/mbbeagle.c:1236-1274);
 *   - TIME_MIN / TIME_MAX special cases of TiProbs_Gen
 *     (src/likelihood.c:9503-9525) instead of BRLENS_MIN clamping
 *     (src/mbbeagle.c:1453-1456);
 *   - every updated non-root interior node is rescaled (rescaleFreq == 1,
 *     src/mcmc.c:6157-6164) with RemoveNodeScalers bookkeeping
 *     (src/likelihood.c:7938-7965).
 *
 * The host keeps ALL integer state: every Flip*Space call the reference makes
 * is made here too, in the same order, so ResetFlips (src/mcmc.c:15695) keeps
 * working unmodified; the engine is addressed purely by the indices valid at
 * call time.  One LogLike call == one mb200_evaluate == one fused GPU pass.
 *
 * No reference source text is copied: only the public structs and function
 * prototypes of the reference headers are used.
 */
#include "bayes.h"
#include "likelihood.h"
#include "mbbeagle.h"
#include "mcmc.h"
#include "model.h"
#include "utils.h"

#include "mb200.h"
#include "mb200_seam.h"

/* globals of src/mcmc.c that have no extern declaration in a header */
extern int *chainId;
extern int  numLocalChains;
/* defined in src/likelihood.c:70, not declared in likelihood.h */
int UpDateCijk (int whichPart, int whichChain);

#define SEAM_MAX_DIVISIONS 512

typedef struct
    {
    int                  instance;          /* engine instance, -1 = none            */
    int                  capOps, capMats;
    mb200_operation     *ops;
    mb200_matrix_update *mats;
    mb200_evaluation     ev;                /* evaluation being assembled            */
    int                  inlineEigen;       /* nst = 1, 2: eigensystem derived per evaluation */
    double               eigenBlock[72];    /* [lambda_re(4), lambda_im(4), c_ijk(64)] */
    long long            clUpdates;         /* node*pattern*rate updates issued      */
    int                  pending;           /* launched by a deferred evaluation, result not yet collected */
    double               syncValue;         /* backend without begin/end: the result, kept until collected */
    int                  syncStatus, syncRc;
    } SeamDivision;

static SeamDivision seamDiv[SEAM_MAX_DIVISIONS];
static int          seamInitialized = NO;
/* cijk slots the device has a copy of (bit per slot).  An evaluation that reads a slot
   the device has never seen uploads it first: chains whose first evaluation ran on the
   reference's own path, or an instance that was re-created. */
static unsigned char seamCijkSeen[SEAM_MAX_DIVISIONS][(MAX_CHAINS + 8) / 8 + 1];

/* ---- backend indirection: lets the oracle harness record or shadow every call ---- */
static int be_create (const mb200_instance_config *c, int *i)          { return mb200_create_instance (c, i); }
static int be_finalize (int i)                                         { return mb200_finalize_instance (i); }
static int be_tips (int i, int t, const uint64_t *m)                   { return mb200_set_tip_states (i, t, m); }
static int be_weights (int i, int r, const float *w)                   { return mb200_set_pattern_weights (i, r, w); }
static int be_cijk (int i, int e, const double *b)                     { return mb200_set_cijk (i, e, b); }
static int be_eval (int i, const mb200_evaluation *e, int n, double *l, int *s) { return mb200_evaluate (i, e, n, l, s); }
static int be_begin (int i, const mb200_evaluation *e, int n)          { return mb200_evaluate_begin (i, e, n); }
static int be_end (int i, double *l, int *s)                           { return mb200_evaluate_end (i, l, s); }

static MB200SeamBackend seamBackend = { be_create, be_finalize, be_tips, be_weights, be_cijk, be_eval, be_begin, be_end };
static int seamDeferred = NO;   /* YES: TreeLikelihood_Beagle only launches; SeamCollect fetches the result */

void MB200SeamSetBackend (const MB200SeamBackend *backend)
{
    if (backend == NULL)
        {
        MB200SeamBackend def = { be_create, be_finalize, be_tips, be_weights, be_cijk, be_eval, be_begin, be_end };
        seamBackend = def;
        }
    else
        seamBackend = *backend;
}

static void SeamInit (void)
{
    int d;
    if (seamInitialized == YES)
        return;
    for (d=0; d<SEAM_MAX_DIVISIONS; d++)
        {
        memset (&seamDiv[d], 0, sizeof(SeamDivision));
        seamDiv[d].instance = -1;
        }
    seamInitialized = YES;
}

long long MB200SeamUpdateCount (int division)
{
    if (seamInitialized == NO || division < 0 || division >= SEAM_MAX_DIVISIONS)
        return 0;
    return seamDiv[division].clUpdates;
}

int MB200SeamInstance (int division)
{
    if (seamInitialized == NO || division < 0 || division >= SEAM_MAX_DIVISIONS)
        return -1;
    return seamDiv[division].instance;
}

/* 4x4 nucleotide models with nst = 1 or 2 (JC69, F81, K80, HKY85): the reference evaluates
 * them with closed forms (TiProbs_JukesCantor / _Fels / _Hky, src/likelihood.c:9289, 9709, 9846)
 * and keeps no eigensystem for them in non-BEAGLE builds (InitEigenSystemInfo, src/mcmc.c:6540-6552).
 * The seam derives the eigensystem of their rate matrix for every evaluation and ships it with
 * the call (mb200_evaluation.inline_eigen), so the engine needs no extra kernel. */
int MB200SeamClosedFormModel (ModelInfo *m)
{
    if ((m->dataType == DNA || m->dataType == RNA) && m->nucModelId == NUCMODEL_4BY4 &&
        (m->nst == 1 || m->nst == 2) && m->numModelStates == 4 && m->switchRates == NULL && m->nCijkParts == 0)
        return YES;
    return NO;
}

/* Q of HKY85 (kappa = 1: F81 / JC69), scaled to one expected substitution per unit time -- the
 * same normalisation as TiProbs_Hky's beta (src/likelihood.c:9745) -- then GetEigens + CalcCijk,
 * the two public utilities UpDateCijk itself uses (src/likelihood.c:10626-10661) */
static int SeamClosedFormEigen (ModelInfo *m, int chain, double *block)
{
    int             i, j, isComplex;
    MrBFlt          kappa, *bs, scaler, mult, **q, **eigvecs, **inverseEigvecs, eigenValues[4], eigvalsImag[4];
    MrBComplex      **Ceigvecs, **CinverseEigvecs;

    bs = GetParamSubVals (m->stateFreq, chain, state[chain]);
    kappa = (m->nst == 2) ? *GetParamVals (m->tRatio, chain, state[chain]) : 1.0;
    q = AllocateSquareDoubleMatrix (4);
    eigvecs = AllocateSquareDoubleMatrix (4);
    inverseEigvecs = AllocateSquareDoubleMatrix (4);
    Ceigvecs = AllocateSquareComplexMatrix (4);
    CinverseEigvecs = AllocateSquareComplexMatrix (4);
    for (i=0; i<4; i++)
        q[i][i] = 0.0;
    scaler = 0.0;
    for (i=0; i<4; i++)
        for (j=i+1; j<4; j++)
            {
            mult = ((i == 0 && j == 2) || (i == 1 && j == 3)) ? kappa : 1.0;   /* A<->G, C<->T */
            q[i][i] -= (q[i][j] = bs[j] * mult);
            q[j][j] -= (q[j][i] = bs[i] * mult);
            scaler += bs[i] * q[i][j];
            scaler += bs[j] * q[j][i];
            }
    scaler = 1.0 / scaler;
    for (i=0; i<4; i++)
        for (j=0; j<4; j++)
            q[i][j] *= scaler;
    isComplex = GetEigens (4, q, eigenValues, eigvalsImag, eigvecs, inverseEigvecs, Ceigvecs, CinverseEigvecs);
    if (isComplex == NO)
        {
        for (i=0; i<4; i++)
            {
            block[i] = eigenValues[i];
            block[4+i] = eigvalsImag[i];
            }
        CalcCijk (4, block + 8, eigvecs, inverseEigvecs);
        }
    FreeSquareDoubleMatrix (q);
    FreeSquareDoubleMatrix (eigvecs);
    FreeSquareDoubleMatrix (inverseEigvecs);
    FreeSquareComplexMatrix (Ceigvecs);
    FreeSquareComplexMatrix (CinverseEigvecs);
    return (isComplex == NO) ? NO_ERROR : ERROR;
}

/* Codon models with omega categories (NY98, M3: lset omegavar=...): one rate matrix, hence one
 * eigensystem, per category (nCijkParts = numOmegaCats, InitEigenSystemInfo src/mcmc.c:6573-6577);
 * P(t) from TiProbs_GenCov (src/likelihood.c:9568), pruning CondLikeDown/Root/Scaler_NY98
 * (:1575, :4010, :5413), root Likelihood_NY98 (:6975): the general-S arithmetic with
 * K = numOmegaCats and the omega category frequencies as category weights. */
static int SeamOmegaCategories (ModelInfo *m)
{
    if ((m->dataType != DNA && m->dataType != RNA) || m->nucModelId != NUCMODEL_CODON)
        return NO;
    if (m->numOmegaCats <= 1 || m->numOmegaCats > MB200_MAX_CATEGORIES || m->numRateCats != 1)
        return NO;
    if (m->nCijkParts != m->numOmegaCats || m->omega == NULL || m->pInvar != NULL)
        return NO;
    return YES;
}

/* Covarion models with gamma rate variation (lset covarion=yes rates=gamma; nucleotide S = 8, protein
 * S = 40): the category's rate sits inside its own rate matrix, so there is one eigensystem per
 * rate category (nCijkParts = numRateCats, InitEigenSystemInfo src/mcmc.c:6533-6563) and P(t) comes
 * from TiProbs_GenCov; everything downstream is the general-S path (CondLike*_Gen*, Likelihood_Gen*).
 * Covarion without rate variation has nCijkParts = 1 and goes through TiProbs_Gen like any other model. */
static int SeamCovarionGamma (ModelInfo *m)
{
    if (m->switchRates == NULL || m->numOmegaCats != 1)
        return NO;
    if (m->nCijkParts <= 1 || m->nCijkParts != m->numRateCats)
        return NO;
    return YES;
}

/* one eigensystem per category? (TiProbs_GenCov models) */
static int SeamCategoryEigens (ModelInfo *m)
{
    return (SeamOmegaCategories (m) == YES || SeamCovarionGamma (m) == YES) ? YES : NO;
}

/* rate / omega categories of the division as the engine sees them */
static int SeamCategories (ModelInfo *m)
{
    return (SeamOmegaCategories (m) == YES) ? m->numOmegaCats : m->numRateCats;
}

/* Which divisions the engine takes; everything else stays on the reference's own
 * function pointers, the way the reference keeps BEAGLE away from models it does
 * not cover (src/mcmc.c:5741-5775). */
int MB200SeamDivisionSupported (ModelInfo *m)
{
    if (m->parsModelId == YES)
        return NO;
    if (m->dataType != DNA && m->dataType != RNA && m->dataType != PROTEIN)
        return NO;                              /* STANDARD / RESTRICTION / CONTINUOUS: next rows */
    if (m->nCijkParts != 1 && MB200SeamClosedFormModel (m) == NO && SeamCategoryEigens (m) == NO)
        return NO;
    if (m->gibbsGamma == YES || m->correlation != NULL)
        return NO;
    if (m->switchRates != NULL)
        return NO;                              /* covarion: the plumbing is here (SeamCovarionGamma, on/off
                                                   frequencies) but the restatement is not pinned: on primates the reference
                                                   (as built here) starts at lnL -1557.87, above anything the data
                                                   allows, while this path and an independent float64 recomputation
                                                   from the same inputs give -8553.72 */
    if (m->numModelStates < 2 || m->numModelStates > MB200_MAX_STATES)
        return NO;
    if (m->numRateCats < 1 || m->numRateCats > MB200_MAX_CATEGORIES)
        return NO;
    if (m->numOmegaCats != 1 && SeamOmegaCategories (m) == NO)
        return NO;
    if (m->nParsIntsPerSite != 1)
        return NO;
    if (m->printAncStates == YES || m->printSiteRates == YES || m->printPosSel == YES || m->printSiteOmegas == YES)
        return NO;                              /* host readers of CL buffers (src/mcmc.c:5761-5772) */
    return YES;
}

/* ---- InitBeagleInstance (src/mbbeagle.c:60): allocate device buffers, load tips ---- */
int InitBeagleInstance (ModelInfo *m, int division)
{
    int                     i, c, b, nRep, rc;
    uint64_t               *masks, obs, full;
    mb200_instance_config   cfg;
    SeamDivision           *sd;

    SeamInit ();
    if (division < 0 || division >= SEAM_MAX_DIVISIONS)
        return (ERROR);
    sd = &seamDiv[division];
    if (sd->instance >= 0)
        return (NO_ERROR);
    if (MB200SeamDivisionSupported (m) == NO)
        return (ERROR);

    memset (&cfg, 0, sizeof(cfg));
    cfg.tip_count       = numLocalTaxa;
    cfg.partials_count  = m->numCondLikes;
    cfg.state_count     = m->numModelStates;
    cfg.pattern_count   = m->numChars;
    cfg.category_count  = SeamCategories (m);
    cfg.flags           = (SeamCategoryEigens (m) == YES) ? MB200_CONFIG_CIJK_PARTS (m->nCijkParts) : 0;
    cfg.matrix_count    = m->numTiProbs;
    cfg.scaler_count    = m->numScalers;
    cfg.eigen_count     = numLocalChains + 1;    /* unused (but harmless) for the inline-eigen models */
    cfg.weight_rows     = chainParams.numChains;
    cfg.device          = 0;
    cfg.max_evaluations = 1;
    rc = seamBackend.create_instance (&cfg, &sd->instance);
    if (rc != MB200_SUCCESS)
        {
        MrBayesPrint ("%s   B200 engine: cannot create instance for division %d (%s)\n", spacer, division+1, mb200_error_string (rc));
        sd->instance = -1;
        return (ERROR);
        }

    /* tip state sets: one bit per model state, hidden-state blocks replicated */
    masks = (uint64_t *) SafeMalloc ((size_t)m->numChars * sizeof(uint64_t));
    if (!masks)
        return (ERROR);
    nRep = m->numModelStates / m->numStates;
    full = (m->numStates == 64) ? ~(uint64_t)0 : (((uint64_t)1 << m->numStates) - 1);
    for (i=0; i<numLocalTaxa; i++)
        {
        for (c=0; c<m->numChars; c++)
            {
            obs = (uint64_t) m->parsSets[i][c * m->nParsIntsPerSite] & full;
            masks[c] = 0;
            for (b=0; b<nRep; b++)
                masks[c] |= obs << (b * m->numStates);
            }
        if (seamBackend.set_tip_states (sd->instance, i, masks) != MB200_SUCCESS)
            {
            free (masks);
            return (ERROR);
            }
        }
    free (masks);

    /* pattern weights, one row per heat-ordered chain id (src/likelihood.c:5830) */
    for (i=0; i<chainParams.numChains; i++)
        {
        if (seamBackend.set_pattern_weights (sd->instance, i, numSitesOfPat + i*numCompressedChars + m->compCharStart) != MB200_SUCCESS)
            return (ERROR);
        }

    sd->capOps  = m->numCondLikes;
    sd->capMats = m->numTiProbs;
    sd->ops  = (mb200_operation *)     SafeCalloc ((size_t)sd->capOps,  sizeof(mb200_operation));
    sd->mats = (mb200_matrix_update *) SafeCalloc ((size_t)sd->capMats, sizeof(mb200_matrix_update));
    if (!sd->ops || !sd->mats)
        return (ERROR);

    MrBayesPrint ("%s   Using B200 engine (%s) for division %d: %d patterns x %d rate cats x %d states\n",
                  spacer, mb200_version_string (), division+1, m->numChars, m->numRateCats, m->numModelStates);
    return (NO_ERROR);
}

void MB200SeamFinalize (void)
{
    int d;
    if (seamInitialized == NO)
        return;
    for (d=0; d<SEAM_MAX_DIVISIONS; d++)
        {
        if (seamDiv[d].instance >= 0)
            seamBackend.finalize_instance (seamDiv[d].instance);
        free (seamDiv[d].ops);
        free (seamDiv[d].mats);
        memset (&seamDiv[d], 0, sizeof(SeamDivision));
        seamDiv[d].instance = -1;
        }
    memset (seamCijkSeen, 0, sizeof(seamCijkSeen));
}

/* branch length seen by the substitution model (src/likelihood.c:9471-9496) */
static MrBFlt SeamBranchLength (ModelInfo *m, TreeNode *p, int chain)
{
    if (m->cppEvents != NULL)
        return GetParamSubVals (m->cppEvents, chain, state[chain])[p->index];
    else if (m->tk02BranchRates != NULL)
        return GetParamSubVals (m->tk02BranchRates, chain, state[chain])[p->index];
    else if (m->wnBranchRates != NULL)
        return GetParamSubVals (m->wnBranchRates, chain, state[chain])[p->index];
    else if (m->ilnBranchRates != NULL)
        return GetParamSubVals (m->ilnBranchRates, chain, state[chain])[p->index];
    else if (m->igrBranchRates != NULL)
        return GetParamSubVals (m->igrBranchRates, chain, state[chain])[p->index];
    else if (m->mixedBrchRates != NULL)
        return GetParamSubVals (m->mixedBrchRates, chain, state[chain])[p->index];
    return p->length;
}

static void SeamQueueMatrix (SeamDivision *sd, ModelInfo *m, TreeNode *p, int chain)
{
    mb200_matrix_update *u;

    FlipTiProbsSpace (m, chain, p->index);
    u = &sd->mats[sd->ev.matrix_update_count++];
    u->matrix = m->tiProbsIndex[chain][p->index];
    u->eigen  = (sd->inlineEigen == YES) ? MB200_EIGEN_INLINE : m->cijkIndex[chain];
    u->length = SeamBranchLength (m, p, chain);
}

/* ---- TreeTiProbs_Beagle (src/mbbeagle.c:1368): which P(t) must be rebuilt ---------- */
int TreeTiProbs_Beagle (Tree *t, int division, int chain)
{
    int             i, k;
    MrBFlt          baseRate, corr, theRate, *catRate, pInvar;
    TreeNode       *p;
    ModelInfo      *m;
    SeamDivision   *sd;

    m  = &modelSettings[division];
    sd = &seamDiv[division];
    sd->ev.matrix_update_count = 0;
    sd->ev.matrix_updates      = sd->mats;

    /* same visiting order and the same flips as src/likelihood.c:7892-7918 */
    for (i=0; i<t->nIntNodes; i++)
        {
        p = t->intDownPass[i];
        if (p->left->upDateTi == YES)
            SeamQueueMatrix (sd, m, p->left, chain);
        if (p->right->upDateTi == YES)
            SeamQueueMatrix (sd, m, p->right, chain);
        if (t->isRooted == NO && p->anc->anc == NULL)
            SeamQueueMatrix (sd, m, p, chain);      /* interior root's branch: always rebuilt */
        }

    /* rate multipliers of TiProbs_Gen (src/likelihood.c:9432-9464) */
    corr = 1.0;
    if (m->dataType == DNA || m->dataType == RNA)
        {
        if (m->nucModelId == NUCMODEL_DOUBLET)
            corr = 2.0;
        else if (m->nucModelId == NUCMODEL_CODON)
            corr = 3.0;
        }
    baseRate = GetRate (division, chain);
    pInvar = 0.0;
    if (m->pInvar != NULL)
        {
        pInvar = *GetParamVals (m->pInvar, chain, state[chain]);
        baseRate /= (1.0 - pInvar);
        }
    theRate = 1.0;
    if (m->shape != NULL)
        catRate = GetParamSubVals (m->shape, chain, state[chain]);
    else if (m->mixtureRates != NULL)
        catRate = GetParamSubVals (m->mixtureRates, chain, state[chain]);
    else
        catRate = &theRate;
    for (k=0; k<m->numRateCats; k++)
        sd->ev.category_rates[k] = baseRate * catRate[k] * corr;
    if (SeamCategoryEigens (m) == YES)
        for (k=0; k<SeamCategories (m); k++)
            sd->ev.category_rates[k] = corr;    /* TiProbs_GenCov: t = length * correctionFactor, nothing else */

    return (NO_ERROR);
}

/* ---- TreeCondLikes_Beagle_Always_Rescale (src/mbbeagle.c:995): the op list --------- */
int TreeCondLikes_Beagle_Always_Rescale (Tree *t, int division, int chain)
{
    int                 i;
    TreeNode           *p;
    ModelInfo          *m;
    SeamDivision       *sd;
    mb200_operation    *op;

    m  = &modelSettings[division];
    sd = &seamDiv[division];
    sd->ev.operation_count = 0;
    sd->ev.operations      = sd->ops;

    for (i=0; i<t->nIntNodes; i++)
        {
        p = t->intDownPass[i];
        if (p->upDateCl != YES)
            continue;

        op = &sd->ops[sd->ev.operation_count++];

        /* CondLikeDown_* / CondLikeRoot_* flip first, then read the child indices
           (src/likelihood.c:795-804) */
        FlipCondLikeSpace (m, chain, p->index);
        op->dest    = m->condLikeIndex[chain][p->index];
        op->child1  = m->condLikeIndex[chain][p->left->index];
        op->matrix1 = m->tiProbsIndex [chain][p->left->index];
        op->child2  = m->condLikeIndex[chain][p->right->index];
        op->matrix2 = m->tiProbsIndex [chain][p->right->index];
        if (t->isRooted == NO && p->anc->anc == NULL)
            {
            op->child3  = m->condLikeIndex[chain][p->anc->index];
            op->matrix3 = m->tiProbsIndex [chain][p->index];
            }
        else
            {
            op->child3  = MB200_NONE;
            op->matrix3 = MB200_NONE;
            }

        /* scaler bookkeeping of src/likelihood.c:7938-7965 */
        if (m->unscaledNodes[chain][p->index] == 0 && m->upDateAll == NO)
            op->scale_remove = m->nodeScalerIndex[chain][p->index];
        else
            op->scale_remove = MB200_NONE;
        FlipNodeScalerSpace (m, chain, p->index);
        m->unscaledNodes[chain][p->index] = 1 + m->unscaledNodes[chain][p->left->index]
                                              + m->unscaledNodes[chain][p->right->index];
        if (m->unscaledNodes[chain][p->index] >= m->rescaleFreq[chain] && p->anc->anc != NULL)
            {
            op->scale_write = m->nodeScalerIndex[chain][p->index];
            m->unscaledNodes[chain][p->index] = 0;
            }
        else
            op->scale_write = MB200_NONE;

        sd->clUpdates += (long long) m->numChars * SeamCategories (m);
        }

    return (NO_ERROR);
}

/* the reference has these two for the dynamic-rescaling scheme; the engine always
   rescales (the built-in path's policy), so both map onto the same op list */
int TreeCondLikes_Beagle_No_Rescale (Tree *t, int division, int chain)
{
    return TreeCondLikes_Beagle_Always_Rescale (t, division, chain);
}

int TreeCondLikes_Beagle_Rescale_All (Tree *t, int division, int chain)
{
    return TreeCondLikes_Beagle_Always_Rescale (t, division, chain);
}

static int SeamApplyResult (int division, int rc, double value, int status, MrBFlt *lnL);

/* ---- TreeLikelihood_Beagle (src/mbbeagle.c:1117): root integration; launches ------- */
int TreeLikelihood_Beagle (Tree *t, int division, int chain, MrBFlt *lnL, int whichSitePats)
{
    int             k, s, status, rc;
    MrBFlt          pInvar, freq, *bs;
    double          value;
    ModelInfo      *m;
    SeamDivision   *sd;

    m  = &modelSettings[division];
    sd = &seamDiv[division];

    sd->ev.root_buffer = m->condLikeIndex[chain][t->root->left->index];
    sd->ev.weights_row = whichSitePats;
    sd->ev.flags       = 0;
    sd->ev.inline_eigen = (sd->inlineEigen == YES) ? sd->eigenBlock : NULL;

    pInvar = 0.0;
    sd->ev.has_p_invar = NO;
    if (m->pInvar != NULL)
        {
        pInvar = *GetParamVals (m->pInvar, chain, state[chain]);
        sd->ev.has_p_invar = YES;
        }
    sd->ev.p_invar = pInvar;
    /* which reference kernel family this division would run (SetLikeFunctions,
       src/mcmc.c:17995-18010 vs 18109-18243) decides two rounding-level details */
    if (m->numModelStates == 4 && (m->dataType == DNA || m->dataType == RNA))
        sd->ev.flags |= MB200_FLAG_NUC4_PINVAR_QUIRK;   /* Likelihood_NUC4_* family */
    else
        sd->ev.flags |= MB200_FLAG_TIP_SHORTCUTS;       /* *_Gen_SSE family */

    /* category weights (src/likelihood.c:5821-5824) */
    if (m->pInvar == NULL)
        freq = 1.0 / m->numRateCats;
    else
        freq = (1.0 - pInvar) / m->numRateCats;
    for (k=0; k<m->numRateCats; k++)
        sd->ev.category_weights[k] = freq;
    if (SeamOmegaCategories (m) == YES)
        {
        /* Likelihood_NY98 (src/likelihood.c:6998): the omega category frequencies */
        MrBFlt *omegaCatFreq = GetParamSubVals (m->omega, chain, state[chain]);
        for (k=0; k<m->numOmegaCats; k++)
            sd->ev.category_weights[k] = omegaCatFreq[k];
        }

    bs = GetParamSubVals (m->stateFreq, chain, state[chain]);
    for (s=0; s<m->numModelStates; s++)
        sd->ev.state_freqs[s] = bs[s];
    if (m->switchRates != NULL)
        {
        /* covarion: stationary frequencies of the on / off copies of every state, on-states first
           (Likelihood_Gen, src/likelihood.c:5799-5818) */
        MrBFlt *swr = GetParamVals (m->switchRates, chain, state[chain]);
        MrBFlt  probOn = swr[0] / (swr[0] + swr[1]), probOff = 1.0 - probOn;
        int     half = m->numModelStates / 2;
        for (s=0; s<half; s++)
            {
            sd->ev.state_freqs[s]        = bs[s] * probOn;
            sd->ev.state_freqs[s + half] = bs[s] * probOff;
            }
        }

    if (seamDeferred == YES)
        {
        /* partition-batched evaluation: launch only, MB200LogLike collects */
        if (seamBackend.evaluate_begin != NULL && seamBackend.evaluate_end != NULL)
            sd->syncRc = seamBackend.evaluate_begin (sd->instance, &sd->ev, 1);
        else
            sd->syncRc = seamBackend.evaluate (sd->instance, &sd->ev, 1, &sd->syncValue, &sd->syncStatus);
        sd->pending = YES;
        return (NO_ERROR);
        }
    rc = seamBackend.evaluate (sd->instance, &sd->ev, 1, &value, &status);
    return SeamApplyResult (division, rc, value, status, lnL);
}

/* result of an evaluation -> the reference's conventions */
static int SeamApplyResult (int division, int rc, double value, int status, MrBFlt *lnL)
{
    if (rc != MB200_SUCCESS)
        {
        MrBayesPrint ("%s   B200 engine: evaluation failed for division %d (%s)\n", spacer, division+1, mb200_error_string (rc));
        (*lnL) = MRBFLT_NEG_MAX;
        abortMove = YES;
        return (ERROR);
        }
    if (status == MB200_EVAL_UNDERFLOW)
        {
        /* same signalling as Likelihood_* (src/likelihood.c:5857-5859) */
        (*lnL) = MRBFLT_NEG_MAX;
        abortMove = YES;
        return (ERROR);
        }
    (*lnL) = value;
    return (NO_ERROR);
}

/* ---- LaunchBEAGLELogLikeForDivision (src/mbbeagle.c:400, ALWAYS scheme) ------------ */
void LaunchBEAGLELogLikeForDivision (int chain, int d, ModelInfo *m, Tree *tree, MrBFlt *lnL)
{
    SeamDivision *sd = &seamDiv[d];

    /* site scalers: flip, then reset or copy (src/likelihood.c:7885-7889); the copy
       itself happens on the device as part of the fused pass */
    FlipSiteScalerSpace (m, chain);
    sd->ev.site_scaler_dst = m->siteScalerIndex[chain];
    sd->ev.site_scaler_src = (m->upDateAll == YES) ? MB200_NONE : m->siteScalerScratchIndex;

    TreeTiProbs_Beagle (tree, d, chain);
    TreeCondLikes_Beagle_Always_Rescale (tree, d, chain);
    TreeLikelihood_Beagle (tree, d, chain, lnL, chainId[chain] % chainParams.numChains);
}

/* ---- replacement for LaunchLogLikeForDivision (src/likelihood.c:7851) -------------- */
/* Returns NO when the division is not handled by the engine (caller keeps the
   reference's own function-pointer path for it). */
int MB200LaunchLogLikeForDivision (int chain, int d, MrBFlt *lnL)
{
    int         idx;
    ModelInfo  *m;
    Tree       *tree;
    SeamDivision *sd;

    SeamInit ();
    m = &modelSettings[d];
    if (MB200SeamDivisionSupported (m) == NO)
        return (NO);
    sd = &seamDiv[d];
    if (sd->instance < 0 && InitBeagleInstance (m, d) == ERROR)
        return (NO);

    tree = GetTree (m->brlens, chain, state[chain]);

    if (MB200SeamClosedFormModel (m) == YES)
        {
        /* no cijk bookkeeping in the reference for these models: derive the eigensystem of the
           chain's current kappa / base frequencies and send it along with the evaluation */
        sd->inlineEigen = YES;
        if (m->upDateCijk == YES)
            m->upDateAll = YES;                 /* what LaunchLogLikeForDivision does (src/likelihood.c:7864-7872) */
        if (SeamClosedFormEigen (m, chain, sd->eigenBlock) == ERROR)
            {
            (*lnL) = MRBFLT_NEG_MAX;
            abortMove = YES;
            return (YES);
            }
        LaunchBEAGLELogLikeForDivision (chain, d, m, tree, lnL);
        return (YES);
        }
    sd->inlineEigen = NO;

    if (m->upDateCijk == YES)
        {
        if (UpDateCijk (d, chain) == ERROR)
            {
            (*lnL) = MRBFLT_NEG_MAX;    /* effectively abort the move */
            return (YES);
            }
        m->upDateAll = YES;
        }
    /* the engine's copy of the eigensystem follows the host's */
    idx = m->cijkIndex[chain];
    if (idx < 0 || idx > MAX_CHAINS)
        return (NO);
    if (m->upDateCijk == YES || (seamCijkSeen[d][idx >> 3] & (1 << (idx & 7))) == 0)
        {
        if (seamBackend.set_cijk (sd->instance, idx, m->cijks[idx]) != MB200_SUCCESS)
            {
            (*lnL) = MRBFLT_NEG_MAX;
            abortMove = YES;
            return (YES);
            }
        seamCijkSeen[d][idx >> 3] |= (unsigned char)(1 << (idx & 7));
        }

    LaunchBEAGLELogLikeForDivision (chain, d, m, tree, lnL);
    return (YES);
}

/* collect the result of a deferred evaluation of division d */
static int SeamCollect (int d, MrBFlt *lnL)
{
    SeamDivision *sd = &seamDiv[d];
    double        value = 0.0;
    int           status = MB200_EVAL_OK, rc = sd->syncRc;

    sd->pending = NO;
    if (seamBackend.evaluate_begin != NULL && seamBackend.evaluate_end != NULL)
        {
        if (rc == MB200_SUCCESS)
            rc = seamBackend.evaluate_end (sd->instance, &value, &status);
        }
    else
        { value = sd->syncValue; status = sd->syncStatus; }
    return SeamApplyResult (d, rc, value, status, lnL);
}

/* ---- replacement for the division loop of LogLike (src/mcmc.c:7421-7441) ------------ */
MrBFlt MB200LogLike (int chain, void (*cpuPath) (int chain, int d, MrBFlt *lnL))
{
    int         d;
    ModelInfo  *m;
    MrBFlt      chainLnLike = 0.0;

    /* pass 1: launch every division that needs updating (engine) or compute it (reference path) */
    seamDeferred = YES;
    for (d=0; d<numCurrentDivisions; d++)
        {
        m = &modelSettings[d];
        if (m->upDateCl != YES)
            continue;
        if (MB200LaunchLogLikeForDivision (chain, d, &(m->lnLike[2*chain + state[chain]])) == NO)
            {
            seamDeferred = NO;
            cpuPath (chain, d, &(m->lnLike[2*chain + state[chain]]));
            seamDeferred = YES;
            }
        }
    seamDeferred = NO;

    /* pass 2: collect */
    for (d=0; d<numCurrentDivisions; d++)
        {
        m = &modelSettings[d];
        if (d < SEAM_MAX_DIVISIONS && seamDiv[d].pending == YES)
            SeamCollect (d, &(m->lnLike[2*chain + state[chain]]));
        }
    if (abortMove == YES)
        return MRBFLT_NEG_MAX;
    for (d=0; d<numCurrentDivisions; d++)
        chainLnLike += modelSettings[d].lnLike[2*chain + state[chain]];
    return chainLnLike;
}

/* ---- LaunchBEAGLELogLikeMultiPartition (src/mbbeagle.h:29, called from
 *      LaunchLogLikeForBeagleMultiPartition, src/likelihood.c:7792-7843, which has already run
 *      UpDateCijk for the divisions it passes): all of them in flight together ---------- */
void LaunchBEAGLELogLikeMultiPartition (int *divisions, int divisionCount, int chain, MrBFlt *lnL)
{
    int         i, d, hadCijk;
    ModelInfo  *m;

    (*lnL) = 0.0;
    seamDeferred = YES;
    for (i=0; i<divisionCount; i++)
        {
        d = divisions[i];
        m = &modelSettings[d];
        hadCijk = m->upDateCijk;
        if (MB200SeamClosedFormModel (m) == NO)
            m->upDateCijk = NO;                 /* the caller flipped and rebuilt the cijk space already ... */
        if (hadCijk == YES && d < SEAM_MAX_DIVISIONS)
            memset (seamCijkSeen[d], 0, sizeof(seamCijkSeen[d]));   /* ... so only the upload is left */
        if (MB200LaunchLogLikeForDivision (chain, d, &(m->lnLike[2*chain + state[chain]])) == NO)
            {
            MrBayesPrint ("%s   B200 engine: division %d is outside the engine's coverage\n", spacer, d+1);
            m->lnLike[2*chain + state[chain]] = MRBFLT_NEG_MAX;
            abortMove = YES;
            }
        m->upDateCijk = hadCijk;
        }
    seamDeferred = NO;
    for (i=0; i<divisionCount; i++)
        {
        d = divisions[i];
        m = &modelSettings[d];
        if (d < SEAM_MAX_DIVISIONS && seamDiv[d].pending == YES)
            SeamCollect (d, &(m->lnLike[2*chain + state[chain]]));
        (*lnL) += m->lnLike[2*chain + state[chain]];
        }
    if (abortMove == YES)
        (*lnL) = MRBFLT_NEG_MAX;
}
