#!/bin/bash
# N-GPU call (4 or 8): parity at the timed shapes, phase traces and a quick bench for both work orders, then the bench line
# with the BASELINE sweeps (every entry with its own parity check)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== parity worker (every launch checked; config B sharded, D4k sampled) world $N"
timeout 400 $TR --master-port 29701 tests/multi_gpu_worker.py 2>&1 | grep -E "RANK|Error|error|Traceback" | sort | tail -60 | tee gpurun_out/r2_parity_w$N.log
echo "=== trace B world $N (default order)"; timeout 150 $TR --master-port 29702 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2_trace_b_w$N.log
echo "=== trace B world $N (FM_LOCAL_FIRST=0: GEMM0 of every source, then GEMM1 remote, GEMM1 local)"
FM_LOCAL_FIRST=0 timeout 150 $TR --master-port 29703 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -14
echo "=== quick bench B, FM_LOCAL_FIRST=0"
FM_LOCAL_FIRST=0 timeout 300 $TR --master-port 29705 bench.py --gpus $N --steps 100 --warmup 16 --sweeps none --no-e2e --no-cpu-baseline --parity-tokens 0 2>&1 | tail -1 | cut -c1-330
echo "=== quick bench B, default order"
timeout 300 $TR --master-port 29706 bench.py --gpus $N --steps 100 --warmup 16 --sweeps none --no-e2e --no-cpu-baseline --parity-tokens 0 2>&1 | tail -1 | cut -c1-330
echo "=== bench B world $N + default sweeps"
FM_BENCH_SWEEP_BUDGET_S=300 FM_BENCH_SWEEP_PARITY_TOKENS=48 timeout 600 $TR --master-port 29704 bench.py --gpus $N --steps 200 --warmup 32 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n$N.json | cut -c1-6000
