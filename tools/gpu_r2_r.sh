#!/bin/bash
# round 2, call R: full GPU test suite + default bench + ncu launch list with the kernels of ce81928
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -14 gpurun_out/pytest_r.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::" gpurun_out/bench_r.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_phase1','kernel_us_per_step','parity_at_scale')})
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['by_kernel'])
p=d['pileup']; print({k:p[k] for k in ('value','e2e','e2e_cli','parity','cpu_baseline')})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_r02_r.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pileup > gpurun_out/ncu_r.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r.log | cut -c1-300
