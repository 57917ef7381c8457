#!/bin/bash
# 2-GPU call: expert-parallel parity (every launch checked), external symmetric memory, phase trace, bench with parity check
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
echo "=== pytest multi (world $N + torch symmetric memory)"
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "parity[$N] or torch_symmetric" 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== trace B (slab output)"; timeout 300 $TR --master-port 29611 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -12
echo "=== trace B (caller tensor output)"; timeout 300 $TR --master-port 29612 scripts/trace_multi.py --cfg B --slab-out 0 2>&1 | grep -v Warning | tail -6
echo "=== trace B no cross pairing"; FM_XPAIR=0 timeout 300 $TR --master-port 29613 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -6
echo "=== trace B gather combine (no fused RED)"; FM_FUSED_COMBINE=0 timeout 300 $TR --master-port 29614 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -6
echo "=== bench B N=$N"; timeout 900 $TR --master-port 29615 bench.py --gpus $N --steps 200 --warmup 32 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n$N.json | cut -c1-1500
