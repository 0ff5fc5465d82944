// mb200_device.cuh -- device-side job format and buffer geometry shared by the kernels
// and the host runtime of the B200 tree-likelihood engine.
//
// Data layout in HBM (per instance == one MrBayes data division on one GPU):
//   tips      uint8  [tip][C]            (S <= 8)  state-set bitmask per pattern
//             uint64 [tip][C]            (always)  same, 64-bit
//   partials  float  [buf][k][c][Sp]     interior conditional likelihoods; Sp = S rounded
//                                        up to a multiple of 4 (61 -> 64) so that every
//                                        (k,c) row is float4-addressable / TMA-legal;
//                                        for S = 4 one (k,c) cell is exactly one float4
//   matrices  float  [mat][k][i][j]      P(t), row = ancestral state (reference layout,
//                                        src/likelihood.c:300-309)
//   scalers   float  [scaler][c]         node scalers and site scalers, one index space
//                                        like m->scalers (src/mcmc.c:6017-6046)
//   eigen     double [slot][2S + S^3]    MrBayes' cijk block (src/likelihood.c:9467)
//   weights   float  [row][c]            numSitesOfPat rows
//   invMask   uint64 [c]                 AND of all tip masks (InitInvCondLikes,
//                                        src/mcmc.c:6631-6800)
#pragma once
#include <stdint.h>

#define MB200_DEV_MAX_CATS   20
#define MB200_DEV_MAX_STATES 64

struct DevEval                      // one LaunchLogLikeForDivision
{
    int    nMat, matOff;            // matrix updates [matOff, matOff+nMat) of the batch
    int    nOp,  opOff;             // operations     [opOff,  opOff+nOp)
    int    siteDst, siteSrc;        // -1: do not store / start from zero
    int    root, weightsRow;        // root -1: no root integration
    int    flags, hasPInvar;
    int    equalWeights, pad0;
    double pInvar;
    double rates  [MB200_DEV_MAX_CATS];
    double catW   [MB200_DEV_MAX_CATS];
    double freqs  [MB200_DEV_MAX_STATES];
};

struct DevMat                       // one P(t) rebuild
{
    int    matrix, eigen;
    double length;
    int    eval;                    // which DevEval supplies rates / freqs
    int    pad[3];
};

struct DevOp                        // one interior-node update
{
    int dest, c1, m1, c2, m2, c3, m3, sw, sr;
    int pad[3];
};

struct DevBatchHeader
{
    int nEval, nMat, nOp, pad;
    // followed by DevEval[nEval], DevMat[nMat], DevOp[nOp] (each 16-byte aligned)
};

struct DevCtx                       // instance geometry + buffer bases, passed by value
{
    int S, Sp, K, C;
    int tipCount, partialsCount, matrixCount, scalerCount, eigenCount, weightRows;
    int tilePatterns;               // patterns per CTA in the evaluation kernels
    int numTiles;
    const uint8_t  *tip8;
    const uint64_t *tip64;
    const int      *tipPartAmbig;   // [tip] 1: some pattern is partially ambiguous (isPartAmbig)
    float          *partials;
    float          *matrices;
    float          *scalers;
    const double   *eigen;
    const float    *weights;
    const uint64_t *invMask;
    double         *tilePartial;    // [maxEval][numTiles] per-tile lnL partial sums
    int            *tileAbort;      // [maxEval][numTiles]
    unsigned int   *ticket;         // [maxEval]
};

static inline size_t mb200_align16 (size_t x) { return (x + 15) & ~(size_t)15; }
