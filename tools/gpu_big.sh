#!/bin/bash
# Full-size large configs (C3/C4/C5) on one GPU (or N via torchrun): bash tools/gpu_big.sh TAG WORKLOAD [N] [extra bench args]
TAG=${1:-big}; WL=${2:-c4}; N=${3:-1}; shift 3
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload $WL --steps 3 --warmup 1 --no-cpu-baseline --verify "$@" > $OUT/bench_${TAG}_$WL.json 2> $OUT/bench_${TAG}_$WL.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload $WL --steps 3 --warmup 1 --no-cpu-baseline --verify "$@" > $OUT/bench_${TAG}_${WL}_n$N.json 2> $OUT/bench_${TAG}_${WL}_n$N.err
fi
echo "bench $WL N=$N exit $?"
F=$OUT/bench_${TAG}_$WL.json; [ "$N" != "1" ] && F=$OUT/bench_${TAG}_${WL}_n$N.json
python - <<PY
import json
for line in open("$F"):
    if not line.startswith("{"): continue
    d=json.loads(line)
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}), "verify", d.get("verify"), d["config"]["workload"], d["roofline"]["kernel_ms_per_step"], "cand", d["config"]["candidates_per_step"], "pos", d["config"]["filter_positives_per_step"], "matched", d["config"]["matched"])
PY
tail -5 ${F%.json}.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | head -3
