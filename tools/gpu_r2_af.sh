#!/bin/bash
# round 2, call AF: the pipeline-driven wait mode (spinning while GPU-bound, sleeping while host-bound) with 16 and 4 host threads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_align_sam.py -m gpu -q -x > gpurun_out/pytest_af.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_af.log | cut -c1-200
for thr in 16 4; do
  BSQ_BENCH_THREADS=$thr BQ_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-pileup --no-cpu-baseline > gpurun_out/bench_af_t$thr.json 2> gpurun_out/bench_af_t$thr.err; echo "threads $thr rc=$?"
  grep "bq_pipeline" gpurun_out/bench_af_t$thr.err | tail -1 | cut -c1-200
  python -c "
import json; d=json.load(open('gpurun_out/bench_af_t$thr.json')); print({k:(d[k]['value'] if isinstance(d[k],dict) else d[k]) for k in ('value','e2e','e2e_phase1')})"
done
