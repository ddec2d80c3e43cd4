#!/bin/bash
# Multi-GPU bench on one box: bash tools/gpu_scale.sh TAG N [extra bench args]
TAG=${1:-scale}; N=${2:-2}; shift 2
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench N=$N exit $?"
python - <<PY
import json
for line in open("$OUT/bench_${TAG}_n$N.json"):
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    print("N", d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}), d["config"]["parallelism"], d["roofline"]["kernel_ms_per_step"])
PY
tail -5 $OUT/bench_${TAG}_n$N.err
