#!/bin/bash
# Multi-GPU session (gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> "<workload:shard[:extra,args[:ENV=1,ENV2=x[:label]]]> ..."
TAG=${1:-r02}
N=${2:-2}
RUNS=${3:-"c2:routed c5:routed"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus_$TAG.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus_$TAG.txt 2>&1
PORT=29511
MODEFLAGS="--verify"
if [ -n "$QUICK" ]; then MODEFLAGS="--no-e2e"; fi
RUNNER=${RUNNER:-}
for run in $RUNS; do
    IFS=: read -r w shard extra envs label <<< "$run"
    extra=${extra//,/ }
    envs=${envs//,/ }
    name=bench_${w}_n${N}_${shard}${label:+_$label}_$TAG
    PORT=$((PORT + 1))
    $RUNNER env $envs timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --workload $w --shard $shard $MODEFLAGS --steps 5 --warmup 2 --no-cpu-baseline $extra > $OUT/$name.json 2> $OUT/$name.err
    echo "== $name exit $?"
    tail -c 3000 $OUT/$name.json; grep -v "^W1\|^\[W\|^$" $OUT/$name.err | tail -8
done
ls -la $OUT | tail -20
