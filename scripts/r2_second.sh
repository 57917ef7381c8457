#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== new GPU tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "aux or dense or relaunch or pipelined or misaligned or C_api or E1" 2>&1 | tail -15
echo "=== pcie"; timeout 120 python scripts/pcie_probe.py 2>&1 | tail -3
echo "=== gantt full"; timeout 300 python scripts/trace_gantt.py --cfg B --label full --save gpurun_out/r2_trace_full.npy 2>&1 | tail -60
for dbg in 1 2; do
echo "=== gantt dbg=$dbg"; FM_DBG_FLAGS=$dbg timeout 300 python scripts/trace_gantt.py --cfg B --label dbg$dbg 2>&1 | tail -45
done
echo "=== gantt dbg=3 (neither) full output"; FM_DBG_FLAGS=3 timeout 300 python scripts/trace_gantt.py --cfg B --label dbg3 2>&1 | tail -60
echo "=== solo CTAs gantt"; FM_PAIR=0 timeout 300 python scripts/trace_gantt.py --cfg B --label solo 2>&1 | tail -45
