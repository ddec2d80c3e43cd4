#!/bin/bash
# BASELINE config 1 at full size through the reference's own CLI (oracle/_ref/PgRC-dev-gpu = unmodified reference objects + the shim):
# synthetic 5 Mbp genome, 4M x 100 bp reads at 0.1 % substitutions; CPU matchers vs GPU matchers behind the same classes, -t 1 (the
# deterministic setting); archives compared byte for byte; stage-4 time = readsMatch[s] of pgrc_res.txt.
OUT=${GRAFT_REPO_ROOT:-$PWD}/gpurun_out
CLI=${GRAFT_REPO_ROOT:-$PWD}/oracle/_ref/PgRC-dev-gpu
mkdir -p $OUT
W=/tmp/cli_c1; rm -rf $W; mkdir -p $W; cd $W
/usr/bin/time -v true 2>/dev/null
python ${GRAFT_REPO_ROOT:-/root/repo}/tools/make_fastq.py $W/in.fastq --genome 5000000 --reads 4000000 --len 100 --err 0.001 --seed 1
ls -la $W/in.fastq
run() {  # tag gpu(0/1) threads flags...
  local tag=$1 gpu=$2 thr=$3; shift 3
  mkdir -p $W/$tag; cd $W/$tag
  local t0=$(date +%s.%N)
  if [ $gpu = 1 ]; then PGRC_GPU_MATCHER=1 $CLI -t $thr "$@" -i $W/in.fastq a.pgrc > log.txt 2>&1; else $CLI -t $thr "$@" -i $W/in.fastq a.pgrc > log.txt 2>&1; fi
  local rc=$? t1=$(date +%s.%N)
  echo "$tag rc=$rc wall=$(python3 -c "print(round($t1 - $t0, 2))") size=$(stat -c %s a.pgrc) $(tail -1 pgrc_res.txt | awk -F'\t' '{print "m="$7" t="$10" total[s]="$12" readsMatch[s]="$16}') $(grep -c "(GPU)" log.txt) gpu-lines; $(grep "Matched .* reads (" log.txt | head -2 | tr '\n' ' ')"
}
{
if [ "$1" != "gpuonly" ]; then run d_cpu 0 1 -s d38; run c_cpu 0 1; run c_cpu_t8 0 8; fi
run d_gpu 1 1 -s d38
run c_gpu 1 1
cmp $W/d_cpu/a.pgrc $W/d_gpu/a.pgrc && echo "mode d: archives identical"
cmp $W/c_cpu/a.pgrc $W/c_gpu/a.pgrc && echo "mode c (default CLI): archives identical"
} 2>&1 | tee $OUT/cli_c1.txt
