#!/bin/bash
# 8-GPU (or N-GPU) call: parity at the timed shapes, phase trace, BASELINE configs C / D / E measured
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== parity worker (every launch checked; config B sharded, D4k sampled) world $N"
timeout 400 $TR --master-port 29701 tests/multi_gpu_worker.py 2>&1 | grep -E "RANK|Error|error|Traceback" | sort | tail -60 | tee gpurun_out/r2_parity_w$N.log
echo "=== trace B world $N"; timeout 150 $TR --master-port 29702 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2_trace_b_w$N.log
if [ "$N" = "8" ]; then
echo "=== bench C world $N (parity on 128 sampled tokens per rank)"
timeout 400 $TR --master-port 29703 bench.py --gpus $N --config C --steps 30 --warmup 5 --parity-tokens 128 --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2_bench_c_n$N.json | cut -c1-1800
SW="C,D1k,D4k,D16k,D64k,E8,E16,E32,E64,E128"
else
SW="E8,E16,E32,E64,E128"
fi
echo "=== bench B world $N + sweeps"
FM_BENCH_SWEEP_BUDGET_S=300 FM_BENCH_SWEEP_PARITY_TOKENS=48 timeout 600 $TR --master-port 29704 bench.py --gpus $N --steps 200 --warmup 32 --sweeps $SW 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n$N.json | cut -c1-6000
