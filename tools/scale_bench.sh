#!/bin/bash
# bench.py under torchrun at N GPUs:  tools/scale_bench.sh N [bench.py args...]
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@"
