#!/bin/bash
# round 2, call F2: full GPU test suite (incl. the boundary tests) + default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_f2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f2.log
tail -16 gpurun_out/pytest_f2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f2.json 2> gpurun_out/bench_f2.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::" gpurun_out/bench_f2.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_f2.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_phase1','kernel_us_per_step','parity_at_scale')})
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['by_kernel'])
p=d['pileup']; print({k:p[k] for k in ('value','e2e','e2e_cli','parity','cpu_baseline')})
PY
