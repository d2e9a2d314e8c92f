#!/bin/bash
# usage (GPU box with N GPUs): tools/scale_bench.sh <tag> "<N list>" "<workload list>" [steps]
tag=$1; ns=$2; wls=$3; steps=${4:-10}
mkdir -p gpurun_out
for w in $wls; do for n in $ns; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py"; fi
  timeout 600 $cmd --gpus $n --workload $w --steps $steps --warmup 3 --no-cpu --no-extra > gpurun_out/scale_${tag}_${w}_$n.json 2> gpurun_out/scale_${tag}_${w}_$n.err
  python - "$w" $n "gpurun_out/scale_${tag}_${w}_$n.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N=" + sys.argv[2], round(d["value"], 2), "Gsites/s ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
done; done
