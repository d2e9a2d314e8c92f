#!/bin/bash
# usage (GPU box): tools/ab_bench.sh <workload> <steps> lib1.so lib2.so ...   — same bench, different builds of the library
w=$1; steps=$2; shift 2
for lib in "$@"; do
  SPK_CUDA_LIB=$PWD/$lib python bench.py --workload $w --steps $steps --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$w', round(d['value'],2), 'Gsites/s  ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))"
done
