#!/bin/bash
# 2-GPU call: peer-write bandwidth probe, single-GPU parity suite (new router paths), then the 2-GPU suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== peer write probe"; timeout 300 scripts/bin/peer_bw_probe 2>&1 | tee gpurun_out/peer_bw_probe.txt
echo "=== pytest gpu parity (1 GPU)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
echo "=== gantt E128 (1 GPU)"
F='^===|per pair|gate_done|disp_prefix|dispatch_end|kernel_end|gate_topk|gate_gemv|disp_rows'
timeout 300 python scripts/trace_gantt.py --cfg E128 --label E128 2>&1 | grep -E "$F"
bash scripts/r2_multi2.sh 2
