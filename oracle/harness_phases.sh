cd "$(dirname "$0")/.."
run() { TMP=$(mktemp -d); sed -e "s/NGEN/$4/" -e "s#OUTPREFIX#$TMP/out#" tests/golden/cmd/$3.nex > $TMP/run.nex
  bin=$1; mode=$2; stem=$3; ngen=$4; shift 4
  env MB200_MODE=$mode MB200_REPORT=$TMP/report.json "$@" timeout 900 oracle/_ref/$bin $TMP/run.nex > $TMP/run.log 2>$TMP/run.err || { tail -5 $TMP/run.log $TMP/run.err; }
  python3 -c "
import json; r=json.loads(open('$TMP/report.json').read().strip().splitlines()[-1]); print('$*', {k:r[k] for k in ('sec_gpu','sec_queue','sec_flush','sec_finish','rescale_retries')})"; rm -rf $TMP; }
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=1
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=1 MB200_RESCALE=dynamic
run mb_b200_batched gpu primates_gtr_g4 20000 MB200_BATCH=1 MB200_RESCALE=dynamic MB200_RESCALE_MAXFREQ=2
