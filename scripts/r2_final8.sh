#!/bin/bash
# 8-GPU call, ordered by importance (the GPU budget may cut it short): the driver's own bench command with the BASELINE
# sweeps and a parity check per entry, then the parity worker (every launch checked), then the phase trace
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench B world $N, the driver's command (--steps 20 --warmup 5), default sweeps, parity on 64 tokens per rank and entry"
FM_BENCH_SWEEP_BUDGET_S=120 FM_BENCH_SWEEP_PARITY_TOKENS=64 timeout 420 $TR --master-port 29704 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n$N.json | cut -c1-1200
echo "=== parity worker (every launch checked; config B sharded, D4k sampled) world $N"
timeout 300 $TR --master-port 29701 tests/multi_gpu_worker.py 2>&1 | grep -E "RANK|Error|error|Traceback" | sort | tail -60 | tee gpurun_out/r2_parity_w$N.log | grep -E "ALL OK|rror|Traceback"
echo "=== trace B world $N"; timeout 120 $TR --master-port 29702 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2_trace_b_w$N.log
