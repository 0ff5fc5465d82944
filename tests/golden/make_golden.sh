#!/bin/bash
# Regenerates tests/golden/*.gold.gz from the UNMODIFIED reference (needs /root/reference;
# run from the repo root after `make -C oracle ref`).  Each file is the evaluation-record
# stream of one short MCMC run of the reference, written by oracle/ref_harness.c in dump mode:
# inputs of every LaunchLogLikeForDivision call + the lnL the reference's own kernels returned.
#   usage: tests/golden/make_golden.sh
set -euo pipefail
cd "$(dirname "$0")/../.."
OUT=tests/golden
TMP=$(mktemp -d)
run () {   # name binary ngen maxevals
    local name=$1 bin=$2 ngen=$3 max=$4 cmd=$5
    sed -e "s/NGEN/$ngen/" -e "s#OUTPREFIX#$TMP/$name#" $OUT/cmd/$cmd.nex > $TMP/$name.nex
    MB200_MODE=dump MB200_DUMP_FILE=$TMP/$name.gold MB200_DUMP_MAX=$max MB200_REPORT=$TMP/$name.json \
        oracle/_ref/$bin $TMP/$name.nex > $TMP/$name.log
    gzip -9 -n -c $TMP/$name.gold > $OUT/$name.gold.gz
    echo "$name: $(cat $TMP/$name.json)"
}
run primates_gtr_g4_fma   mb_b200     60 400 primates_gtr_g4
run primates_gtr_g4_sse   mb_b200_sse 30 200 primates_gtr_g4
run primates_gtr_ig4_fma  mb_b200    100 200 primates_gtr_ig4
run primates_gtr_eq_fma   mb_b200    100 150 primates_gtr_eq
run ovomucoids_wag_g4_sse mb_b200    100  80 ovomucoids_wag_g4
run replicase_m0_sse      mb_b200    100  40 replicase_m0
run primates_hky_g4_fma   mb_b200    100 200 primates_hky_g4
run primates_f81_i_fma    mb_b200    100 150 primates_f81_i
MB200_NO_STD=1 run cynmix_part_fma       mb_b200     40 160 cynmix_part    # DNA partitions only (morphology left on the CPU)
run replicase_ny98_sse    mb_b200     60  40 replicase_ny98
run cynmix_full_fma       mb_b200     40 200 cynmix_full
rm -rf $TMP
ls -la $OUT/*.gold.gz
