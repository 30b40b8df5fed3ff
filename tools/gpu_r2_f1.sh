#!/bin/bash
# round 2, call F1 (2 GPUs): the two-GPU product tests, and bench.py under torchrun at N=2 (short)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus2.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus" > gpurun_out/pytest_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f1.log
tail -5 gpurun_out/pytest_f1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_f1_n2.json 2> gpurun_out/bench_f1_n2.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::" gpurun_out/bench_f1_n2.err | tail -12
tail -c 2500 gpurun_out/bench_f1_n2.json
