#!/bin/bash
# round 2, call AC: blocking waits (event with cudaEventBlockingSync) against spinning cudaStreamSynchronize, end to end,
# with 16 and with 4 host threads; align parity tests on the blocking build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_sam.py tests/test_dp.py tests/test_phase1.py -m gpu -q -x > gpurun_out/pytest_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ac.log
tail -3 gpurun_out/pytest_ac.log | cut -c1-300
for thr in 16 4; do for spin in 0 1; do
  BSQ_BENCH_THREADS=$thr BSQ_SPIN_WAIT=$spin BQ_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pileup --no-cpu-baseline > gpurun_out/bench_ac_t${thr}_s$spin.json 2> gpurun_out/bench_ac_t${thr}_s$spin.err; echo "threads $thr spin $spin rc=$?"
  grep "bq_pipeline" gpurun_out/bench_ac_t${thr}_s$spin.err | tail -1 | cut -c1-200
  python -c "
import json; d=json.load(open('gpurun_out/bench_ac_t${thr}_s$spin.json')); print({k:(d[k]['value'] if isinstance(d[k],dict) else d[k]) for k in ('value','e2e','e2e_phase1')})"
done; done
