#!/usr/bin/env python3
"""TEST INFRASTRUCTURE / build-time tool: cut RunChain's chain loop in two around LogLike.

usage: patch_runchain.py <reference>/src/mcmc.c <out>/mcmc.c

The reference's generation loop (src/mcmc.c:16718-16938) proposes, evaluates and accepts or rejects ONE chain at
a time, so an accelerator sees one evaluation per call.  Nothing inside a generation couples the chains except
the random-number stream, and LogLike draws no random numbers; the loop can therefore be cut at the LogLike call:

    phase A  for every local chain: copy state, pick and make the move, QUEUE the likelihood evaluation,
             draw the chain's acceptance variate (same position in the stream as in the serial loop)
    flush    ONE engine call per division for all queued evaluations
    phase B  for every local chain: fetch lnL, acceptance probability, accept / reject (ResetFlips), autotune

This script writes a patched COPY of mcmc.c (four anchored text edits, each asserted to match exactly once); no
reference source is stored in the repository, the copy lives under oracle/_ref/ with the other build outputs.
The hooks (MB200RC_*) are defined by the harness (oracle/ref_harness.c) on top of the seam's MB200Batch* API;
a generation the seam cannot batch (MB200RC_Begin() == NO: batching off, a division outside the engine) runs the
ORIGINAL loop body, one step per chain -- the patched binary then reproduces the serial one bit for bit, which is
the self-check tests/test_seam_harness.py runs.  INTEGRATION.md shows the same edit as a maintainer's diff.
"""
import sys

DECLS = r'''
/* ---- chain-batched generations (added by oracle/patch_runchain.py) ---- */
typedef struct { MCMCMove *theMove; MrBFlt lnPriorRatio, lnProposalRatio, u; int abortMove; } MB200RCSlot;
static MB200RCSlot mb200rc[MAX_CHAINS];
static int mb200rcBatch, mb200rcStep;
int    MB200RC_patched = 1;
int    MB200RC_Begin (void);
void   MB200RC_Enter (int chain, int phase);
void   MB200RC_Leave (int chain, int phase);
void   MB200RC_Queue (int chain);
void   MB200RC_Flush (void);
MrBFlt MB200RC_Finish (int chain);

'''

# The chain loop becomes a loop over STEPS.  Serial generation (MB200RC_Begin() == NO): one step per chain, the
# original body untouched (LogLike called where it always was).  Batched generation: steps 0 .. n-1 run phase A of
# chain 0 .. n-1, the queue is flushed, steps n .. 2n-1 run phase B.
EDITS = [
    # (1) loop header; phase A opens: the chain owns its scratch sets while it is being worked on
    ('''        for (chn=0; chn<numLocalChains; chn++)
            {
            /* Do Gibbs resampling of rate categories for current state if time to do so */''',
     '''        mb200rcBatch = MB200RC_Begin ();
        for (mb200rcStep=0; mb200rcStep<(mb200rcBatch == YES ? 2 : 1)*numLocalChains; mb200rcStep++)
            {
            chn = mb200rcStep % numLocalChains;
            if (mb200rcBatch == YES && mb200rcStep == numLocalChains)
                MB200RC_Flush ();
            if (mb200rcBatch == NO || mb200rcStep < numLocalChains)
            {
            if (mb200rcBatch == YES)
                MB200RC_Enter (chn, 0);
            /* Do Gibbs resampling of rate categories for current state if time to do so */'''),
    # (2) the cut: serial = evaluate here; batched = queue, draw the acceptance variate, close phase A / open phase B
    ('''            if (abortMove == NO)
                lnLike = LogLike(chn);
''',
     '''            if (mb200rcBatch == NO)
                {
                if (abortMove == NO)
                    lnLike = LogLike(chn);
                }
            else
                {
                mb200rc[chn].theMove = theMove;
                mb200rc[chn].lnPriorRatio = lnPriorRatio;
                mb200rc[chn].lnProposalRatio = lnProposalRatio;
                if (abortMove == NO)
                    MB200RC_Queue (chn);
                mb200rc[chn].abortMove = abortMove;
                mb200rc[chn].u = (abortMove == NO) ? RandomNumber(seed) : 0.0;
                MB200RC_Leave (chn, 0);
                }
            }
            if (mb200rcBatch == NO || mb200rcStep >= numLocalChains)
            {
            if (mb200rcBatch == YES)
                {
                MB200RC_Enter (chn, 1);
                theMove = mb200rc[chn].theMove;
                lnPriorRatio = mb200rc[chn].lnPriorRatio;
                lnProposalRatio = mb200rc[chn].lnProposalRatio;
                abortMove = mb200rc[chn].abortMove;
                if (abortMove == NO)
                    lnLike = MB200RC_Finish (chn);
                }
'''),
    # (3) the acceptance test: the variate drawn in phase A, or the serial draw
    ('''            if (abortMove == NO && RandomNumber(seed) < r)''',
     '''            if (abortMove == NO && (mb200rcBatch == YES ? mb200rc[chn].u : RandomNumber(seed)) < r)'''),
    # (4) phase B closes
    ('''                maxLnL0[chainId[chn]] = curLnL[chn];
''',
     '''                maxLnL0[chainId[chn]] = curLnL[chn];
            if (mb200rcBatch == YES)
                MB200RC_Leave (chn, 1);
            }
'''),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    text = open(src, encoding="latin-1").read()
    for old, new in EDITS:
        if text.count(old) != 1:
            sys.exit(f"patch_runchain: anchor matches {text.count(old)} times (expected 1):\n{old}")
        text = text.replace(old, new)
    anchor = "int RunChain (RandLong *seed)\n"
    if text.count(anchor) != 1:
        sys.exit("patch_runchain: RunChain definition not found exactly once")
    text = text.replace(anchor, DECLS + anchor)
    open(dst, "w", encoding="latin-1").write(text)


if __name__ == "__main__":
    main()
