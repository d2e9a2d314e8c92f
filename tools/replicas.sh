#!/bin/bash
# usage: tools/replicas.sh <workload> <ngpus> [steps] — N independent single-GPU benches at the same time (no exchange)
w=$1; n=$2; steps=${3:-10}
for i in $(seq 0 $((n-1))); do
  CUDA_VISIBLE_DEVICES=$i python bench.py --workload $w --steps $steps --warmup 3 --no-cpu --no-extra > gpurun_out/replica_${w}_$i.json 2>/dev/null &
done
wait
python - "$w" $n <<'PY'
import json, sys
w, n = sys.argv[1], int(sys.argv[2])
v = [json.loads(open("gpurun_out/replica_%s_%d.json" % (w, i)).read().strip().splitlines()[-1])["value"] for i in range(n)]
print("replicas", w, [round(x, 1) for x in v], "sum", round(sum(v), 1), "min*n", round(min(v) * n, 1))
PY
