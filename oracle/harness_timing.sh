#!/bin/bash
# in-kernel and wall time of the reference-driven runs: cpu vs gpu (per-chain launches) vs gpu chain-batched
cd "$(dirname "$0")/.."
run() { TMP=$(mktemp -d); sed -e "s/NGEN/$4/" -e "s#OUTPREFIX#$TMP/out#" tests/golden/cmd/$3.nex > $TMP/run.nex
  bin=$1; mode=$2; stem=$3; ngen=$4; shift 4
  S=$(date +%s.%N)
  env MB200_MODE=$mode MB200_REPORT=$TMP/report.json "$@" timeout 900 oracle/_ref/$bin $TMP/run.nex > $TMP/run.log 2>$TMP/run.err || { tail -5 $TMP/run.log $TMP/run.err; }
  E=$(date +%s.%N)
  python3 -c "
import json; r=json.loads(open('$TMP/report.json').read().strip().splitlines()[-1]); r['wall_s']=$E-$S; r['binary']='$bin'; r['workload']='$stem'; r['ngen']=$ngen; r['env']='$*'; print(json.dumps(r))"; rm -rf $TMP; }
run mb_b200 cpu primates_gtr_g4 20000
run mb_b200 gpu primates_gtr_g4 20000
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=0
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=1
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=1 MB200_RESCALE=dynamic
run mb_b200 cpu cynmix_full 2000
run mb_b200_batched gpu cynmix_full 2000 MB200_BATCH=0
run mb_b200_batched gpu cynmix_full 2000 MB200_BATCH=1
run mb_b200_batched gpu cynmix_full 2000 MB200_BATCH=1 MB200_RESCALE=dynamic
run mb_b200 cpu replicase_ny98 1000
run mb_b200_batched gpu replicase_ny98 1000 MB200_BATCH=1 MB200_EIGEN=host
run mb_b200_batched gpu replicase_ny98 1000 MB200_BATCH=1 MB200_EIGEN=device
run mb_b200_batched gpu replicase_m0 2000 MB200_BATCH=1 MB200_EIGEN=host
run mb_b200_batched gpu replicase_m0 2000 MB200_BATCH=1 MB200_EIGEN=device
