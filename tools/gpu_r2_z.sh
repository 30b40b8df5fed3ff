#!/bin/bash
# round 2, call Z (4 GPUs): bench.py under torchrun at N=4 with the final code (host memory guard of the pileup leg, e2e with 1/4 of the host threads per rank)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus4_z.txt; nproc >> gpurun_out/gpus4_z.txt; free -g | head -2 >> gpurun_out/gpus4_z.txt
BQ_TIMING=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_z_n4.json 2> gpurun_out/bench_z_n4.err; echo "bench rc=$?"
cat gpurun_out/gpus4_z.txt
grep "bq_pipeline" gpurun_out/bench_z_n4.err | tail -4 | cut -c1-200
grep "bq_finish_a\|bq_finish_b" gpurun_out/bench_z_n4.err | tail -4 | cut -c1-200
grep "pileup rank 0\|shrunk\|failed" gpurun_out/bench_z_n4.err | tail -5 | cut -c1-250
python -c "
import json; d=json.load(open('gpurun_out/bench_z_n4.json')); print({k:d[k] for k in ('value','n_gpus','e2e','e2e_phase1','clocks')}); p=d['pileup']; print({k:p.get(k) for k in ('value','e2e','stats_reduce_ms','n_gpus','contig_shrunk_to_fit_host_memory','error')})"
