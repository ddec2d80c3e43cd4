#!/bin/bash
# A BASELINE config through the reference's own CLI on the GPU box: the same FASTQ compressed by oracle/_ref/PgRC-dev-gpu
# (unmodified reference objects + the pgrc_b200/host shim) with the CPU matchers and with the GPU matchers (stage 4 and
# stage 7), at -t 1 (the reference's deterministic arm), modes d and c (the default CLI).  Archives are compared with cmp;
# stage times come from pgrc_res.txt and the log.
#   bash tools/gpu_cli_fullsize.sh <tag> <name> <genome> <reads> <len> <err> "<extra PgRC flags>"
TAG=$1; NAME=$2; GENOME=$3; READS=$4; LEN=$5; ERR=$6; FLAGS=$7
OUT=$PWD/gpurun_out
CLI=$PWD/oracle/_ref/PgRC-dev-gpu
W=/tmp/cli_$NAME
rm -rf $W; mkdir -p $W $OUT
REP=$OUT/cli_${NAME}_$TAG.txt
echo "config $NAME through the CLI: genome $GENOME, $READS reads x $LEN bp, error rate $ERR, flags '$FLAGS' ($(nproc) host cores)" > $REP
python tools/make_fastq.py $W/in.fastq --genome $GENOME --reads $READS --len $LEN --err $ERR --seed 20261017 || exit 1
ls -l $W/in.fastq >> $REP
run() {   # <arm> <mode flags> <env ...>
    local arm=$1 mode=$2; shift 2
    mkdir -p $W/$arm; cd $W/$arm
    env "$@" $CLI -t 1 $FLAGS $mode -i $W/in.fastq a.pgrc > log.txt 2>&1; local rc=$?
    cd - > /dev/null
    echo "== $arm (exit $rc): $(stat -c %s $W/$arm/a.pgrc 2>/dev/null) bytes" >> $REP
    tail -1 $W/$arm/pgrc_res.txt | awk -F'\t' '{print "   total[s] " $12 "  div " $13 "  PgDiv " $14 "  good " $15 "  readsMatch[s] " $16 "  bad&N " $17 "  order " $18 "  pgSeq-s[s] " $19}' >> $REP
    grep -E "exact matches in|Matched .* reads in|Feeding reference|PgMatching|text index on the GPU" $W/$arm/log.txt | sed 's/^/   /' >> $REP
}
run d_cpu "-s d38" PGRC_GPU_MATCHER=0
run d_gpu "-s d38" PGRC_GPU_MATCHER=1
run c_cpu "" PGRC_GPU_MATCHER=0
run c_gpu "" PGRC_GPU_MATCHER=1
cmp $W/d_cpu/a.pgrc $W/d_gpu/a.pgrc && echo "mode d: archives identical (cmp)" >> $REP || echo "mode d: ARCHIVES DIFFER" >> $REP
cmp $W/c_cpu/a.pgrc $W/c_gpu/a.pgrc && echo "default CLI (mode c): archives identical (cmp)" >> $REP || echo "mode c: ARCHIVES DIFFER" >> $REP
cat $REP
rm -rf $W
