#!/bin/bash
# round 2, call AA: where the host side of a batch goes with 4 threads per GPU (what a rank of an 8-GPU node has), N=1
mkdir -p gpurun_out
BSQ_BENCH_THREADS=4 BQ_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pileup --no-cpu-baseline > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err; echo "bench rc=$?"
grep "bq_pipeline" gpurun_out/bench_aa.err | tail -1 | cut -c1-300
grep "bq_finish_a\|bq_finish_b" gpurun_out/bench_aa.err | tail -8 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_aa.json')); print({k:d[k] for k in ('value','e2e','e2e_phase1')})"
