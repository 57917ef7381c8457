#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu (all single-GPU files; tensor-core router default)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
F='^===|per pair|gate_done|disp_prefix|dispatch_end|ffn_end|kernel_end|barrier|gate_topk|gate_gemv|disp_rows'
echo "=== gantt B tc gate"; timeout 300 python scripts/trace_gantt.py --cfg B --label tc 2>&1 | grep -E "$F"
echo "=== gantt B cuda-core gate"; FM_TC_GATE=0 timeout 300 python scripts/trace_gantt.py --cfg B --label gemv 2>&1 | grep -E "$F"
echo "=== gantt E128 (1 GPU) tc"; timeout 300 python scripts/trace_gantt.py --cfg E128 --label E128tc 2>&1 | grep -E "$F"
echo "=== gantt E128 (1 GPU) gemv"; FM_TC_GATE=0 timeout 300 python scripts/trace_gantt.py --cfg E128 --label E128gemv 2>&1 | grep -E "$F"
echo "=== gantt D64k (1 GPU) tc"; timeout 300 python scripts/trace_gantt.py --cfg D64k --label D64ktc 2>&1 | grep -E "$F"
echo "=== bench B"; timeout 600 python bench.py --steps 200 --warmup 32 2>&1 | tail -1 | tee gpurun_out/r2_bench_b4.json | cut -c1-300
