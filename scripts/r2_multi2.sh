#!/bin/bash
# 2-GPU call: expert-parallel parity (every launch checked), external symmetric memory, phase trace, bench with parity check
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
echo "=== pytest multi (world $N + torch symmetric memory)"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "parity[$N] or torch_symmetric" 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== trace B (slab output)"; timeout 300 $TR --master-port 29611 scripts/trace_multi.py --cfg B --slab-out 1 2>&1 | grep -v Warning | tail -12
echo "=== bench B N=$N"; timeout 900 $TR --master-port 29615 bench.py --gpus $N --steps 200 --warmup 32 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n$N.json | cut -c1-1500
[ -n "$FM_SINGLE_TOO" ] || exit 0
echo "=== single GPU on this box: gantt + bench"
F='^===|per pair|gate_done|disp_prefix|dispatch_end|ffn_end|kernel_end|barrier|gate_topk|gate_gemv|disp_rows'
timeout 300 python scripts/trace_gantt.py --cfg B --label default 2>&1 | grep -E "$F"
timeout 600 python bench.py --steps 200 --warmup 32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_b3.json | cut -c1-300
