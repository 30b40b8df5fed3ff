#!/bin/bash
# round 2, call AE (8 GPUs): bench.py under torchrun at N=8 (the driver's largest scaling point): host memory guard, 4 host threads per rank
mkdir -p gpurun_out
nproc > gpurun_out/box8_ae.txt; free -g | head -2 >> gpurun_out/box8_ae.txt
BQ_TIMING=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_ae_n8.json 2> gpurun_out/bench_ae_n8.err; echo "bench rc=$?"
cat gpurun_out/box8_ae.txt
grep "bq_pipeline" gpurun_out/bench_ae_n8.err | tail -3 | cut -c1-200
grep "shrunk\|failed\|Error\|error" gpurun_out/bench_ae_n8.err | tail -5 | cut -c1-250
python -c "
import json; d=json.load(open('gpurun_out/bench_ae_n8.json')); print({k:d[k] for k in ('value','n_gpus','e2e','e2e_phase1','clocks')}); p=d['pileup']; print({k:p.get(k) for k in ('value','e2e','stats_reduce_ms','n_gpus','contig_shrunk_to_fit_host_memory','error')})"
