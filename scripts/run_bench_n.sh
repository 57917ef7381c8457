#!/bin/bash
# usage: scripts/run_bench_n.sh N [steps] [warmup]   -- the driver's multi-GPU launch line
N=${1:-2}; K=${2:-100}; W=${3:-20}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $K --warmup $W
