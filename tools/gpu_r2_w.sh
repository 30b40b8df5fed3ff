#!/bin/bash
# round 2, call W: full GPU suite + full default bench (align + pileup legs) on the final code
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_w.log
tail -10 gpurun_out/pytest_w.log | cut -c1-300
BQ_TIMING=1 timeout 1500 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc=$?"
grep "bq_pipeline\|parity_at_scale\|\[pileup\]" gpurun_out/bench_w.err | tail -4 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_w.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_phase1','phase2_dp','clocks','gpu_launches')}); r=d['roofline']; print(r['kernel'][:30], r['frac'], r['traffic'], r['moved_layout']['frac'], r['step']['frac']); print(d['cpu_baseline'])
p=d['pileup']; print({k:p[k] for k in ('value','e2e','e2e_cli','parity','cpu_baseline','clocks')}); print(p['roofline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_w_ref.json 2> gpurun_out/bench_w_ref.err; echo "ref arm rc=$?"; cut -c1-900 gpurun_out/bench_w_ref.json
