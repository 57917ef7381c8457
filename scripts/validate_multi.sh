#!/bin/bash
# usage: scripts/validate_multi.sh N  -- expert-parallel parity on N GPUs, then the bench line at N (driver's launch line)
N=${1:-8}
export FM_MULTI_CASES='[{"num_experts":8,"expert_top_k":2,"sequence_len":512,"hidden_size":256,"intermediate_size":512},{"num_experts":16,"expert_top_k":2,"sequence_len":1024,"hidden_size":512,"intermediate_size":1024,"drop_tokens":0},{"num_experts":32,"expert_top_k":4,"sequence_len":256,"hidden_size":128,"intermediate_size":256,"hidden_act":1}]'
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py 2>&1 | grep -E "RANK|Error|error" | sort | tail -40
unset FM_MULTI_CASES
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 100 --warmup 20 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-420
