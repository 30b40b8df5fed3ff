#!/bin/bash
# round 2, call X (2 GPUs): the two-GPU product tests (align -G 0-1 with a DP context per lane, per-rank pileup) and
# bench.py under torchrun at N=2 with the final code
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus2_x.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus or two_ranks or multirank" > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_x.log
tail -5 gpurun_out/pytest_x.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_x_n2.json 2> gpurun_out/bench_x_n2.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::" gpurun_out/bench_x_n2.err | tail -6 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_x_n2.json')); print({k:d[k] for k in ('value','n_gpus','e2e','e2e_phase1','phase2_dp','clocks')}); p=d['pileup']; print({k:p.get(k) for k in ('value','e2e','stats_reduce_ms','n_gpus')})"
