#!/bin/bash
# round 2, call AJ: wait modes on one box with 4 host threads: spinning, sleeping, pipeline-driven
mkdir -p gpurun_out
for mode in spin block adaptive; do
  unset BSQ_SPIN_WAIT BQ_ADAPTIVE_WAIT
  if [ $mode = spin ]; then export BSQ_SPIN_WAIT=1; fi
  if [ $mode = block ]; then export BSQ_SPIN_WAIT=0; fi
  if [ $mode = adaptive ]; then export BQ_ADAPTIVE_WAIT=1; fi
  BSQ_BENCH_THREADS=4 BQ_TIMING=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-pileup --no-cpu-baseline --no-index-check > gpurun_out/bench_aj_$mode.json 2> gpurun_out/bench_aj_$mode.err; echo "$mode rc=$?"
  grep "bq_pipeline" gpurun_out/bench_aj_$mode.err | tail -1 | cut -c1-200
  python -c "
import json; d=json.load(open('gpurun_out/bench_aj_$mode.json')); print('$mode', d['e2e']['value'])"
done
