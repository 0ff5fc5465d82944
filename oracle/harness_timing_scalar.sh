#!/bin/bash
# in-kernel time of the runs whose reference path is the scalar kernel family (no-SIMD build): cpu vs engine
cd "$(dirname "$0")/.."
run() { TMP=$(mktemp -d); sed -e "s/NGEN/$4/" -e "s#OUTPREFIX#$TMP/out#" tests/golden/cmd/$3.nex > $TMP/run.nex
  bin=$1; mode=$2; stem=$3; ngen=$4; shift 4
  S=$(date +%s.%N)
  env MB200_MODE=$mode MB200_REPORT=$TMP/report.json "$@" timeout 900 oracle/_ref/$bin $TMP/run.nex > $TMP/run.log 2>$TMP/run.err || { tail -5 $TMP/run.log $TMP/run.err; }
  E=$(date +%s.%N)
  python3 -c "
import json; r=json.loads(open('$TMP/report.json').read().strip().splitlines()[-1]); r['wall_s']=$E-$S; r['binary']='$bin'; r['workload']='$stem'; r['ngen']=$ngen; r['env']='$*'; print(json.dumps(r))"; rm -rf $TMP; }
run mb_b200_scalar cpu kim_mixed 2000
run mb_b200_scalar gpu kim_mixed 2000
run mb_b200_scalar cpu primates_covarion 5000
run mb_b200_scalar gpu primates_covarion 5000
run mb_b200_scalar cpu ovomucoids_covarion 300
run mb_b200_scalar gpu ovomucoids_covarion 300
run mb_b200_scalar cpu hymfossil_te 2000
run mb_b200_scalar gpu hymfossil_te 2000
# chain-batched generations in the same build
run mb_b200_scalar_batched gpu kim_mixed 2000 MB200_BATCH=1
run mb_b200_scalar_batched gpu primates_covarion 5000 MB200_BATCH=1
run mb_b200_scalar_batched gpu ovomucoids_covarion 300 MB200_BATCH=1
run mb_b200_scalar_batched gpu hymfossil_te 2000 MB200_BATCH=1
