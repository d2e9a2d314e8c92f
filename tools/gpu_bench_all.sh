#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_bench_all.sh <tag> [steps]
# runs bench.py on the four BASELINE workloads and prints one summary line each
tag=${1:-run}; steps=${2:-10}
mkdir -p gpurun_out
for w in potts_neighonly_sc26n_1024 potts_pin_sc26n_1024 ising_sq4n_16384 potts_tri_32768; do
  python bench.py --workload $w --steps $steps --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  python - "$w" "gpurun_out/bench_${tag}_$w.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], round(d["value"], 2), "Gsites/s frac", round(d["roofline"]["frac"], 3), "ms/step",
          round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "acc", round(d["accept_ratio"], 3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 gpurun_out/bench_${tag}_$w.err
done
