#!/bin/bash
# compute-sanitizer memcheck over engine-driven runs of the real program (all kernel families in one process)
cd "$(dirname "$0")/.."
for stem in kim_mixed hymfossil_te primates_covarion replicase_possel cynmix_ordered; do
  TMP=$(mktemp -d); sed -e "s/NGEN/40/" -e "s#OUTPREFIX#$TMP/out#" tests/golden/cmd/$stem.nex > $TMP/run.nex
  bin=mb_b200_scalar_batched
  env MB200_MODE=gpu MB200_BATCH=1 MB200_REPORT=$TMP/r.json timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 oracle/_ref/$bin $TMP/run.nex > $TMP/log.txt 2>&1
  echo "$stem rc=$? $(grep 'ERROR SUMMARY' $TMP/log.txt) $(python3 -c "import json; r=json.loads(open('$TMP/r.json').read().strip().splitlines()[-1]); print('calls', r['calls'], 'unsupported', r['unsupported_calls'], 'batched', r['batched_generations'])" 2>/dev/null)"
  rm -rf $TMP
done
