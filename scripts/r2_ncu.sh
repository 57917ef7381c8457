#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench command + one full capture of the fused kernel; then the
# refined parity check on the tensor-core router (128 experts, 256 tokens) and a memcheck of the smoke invocation
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --parity-tokens 0"
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 20 --csv --log-file gpurun_out/r02_launches_configB.csv $BENCH > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_launches_configB.csv | cut -c1-300
echo "=== full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fm_moe_forward -s 8 -c 2 -f -o gpurun_out/r02_prof $BENCH > gpurun_out/r02_bench_under_ncu_full.log 2>&1
ls -la gpurun_out/r02_prof.ncu-rep
echo "=== parity check, 128 experts on one GPU (tensor-core router), 256 tokens"
timeout 300 python bench.py --config E128 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --parity-tokens 256 2>&1 | tail -1 | tee gpurun_out/r2_bench_e128_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], json.dumps(d['parity_check'])[:900])"
echo "=== compute-sanitizer memcheck on the smoke invocation (small shape)"
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -8 | tee gpurun_out/r02_memcheck_smoke.txt
