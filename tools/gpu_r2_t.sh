#!/bin/bash
# round 2, call T: full GPU suite and default bench with the phase-2 DP kernels in the product path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_t.log
tail -12 gpurun_out/pytest_t.log | cut -c1-300
BQ_TIMING=1 timeout 1200 python bench.py --steps 10 --warmup 3 --no-pileup > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::mem\|bq_batch_run" gpurun_out/bench_t.err | tail -40 | cut -c1-260
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_t.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_phase1','kernel_us_per_step','parity_at_scale')})
PY
