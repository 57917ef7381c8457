#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_large_shapes.py -x -q -m gpu 2>&1 | tail -6
F='^===|per pair|gate_done|disp_prefix|dispatch_end|kernel_end|gate_topk|gate_gemv|disp_rows'
echo "=== gantt B default (gemv)"; timeout 300 python scripts/trace_gantt.py --cfg B --label default 2>&1 | grep -E "$F"
echo "=== gantt B tc gate"; FM_TC_GATE=1 timeout 300 python scripts/trace_gantt.py --cfg B --label tc 2>&1 | grep -E "$F"
echo "=== gantt E128 (1 GPU)"; timeout 300 python scripts/trace_gantt.py --cfg E128 --label E128 2>&1 | grep -E "$F"
echo "=== gantt D16k (1 GPU)"; timeout 300 python scripts/trace_gantt.py --cfg D16k --label D16k 2>&1 | grep -E "$F"
echo "=== bench B"; timeout 600 python bench.py --steps 200 --warmup 32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_b5.json | cut -c1-300
