#!/bin/bash
# round 2, call AD: final validation on one GPU: smoke(), full GPU suite, default bench, reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ad.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_ad.log | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_ad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ad.log
tail -4 gpurun_out/pytest_ad.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ad.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_phase1','clocks','gpu_launches')}); r=d['roofline']; print(r['kernel'][:30], r['frac'], r['traffic'], r['moved_layout']['frac'], r['step']['frac']); print(d['parity_at_scale']['identical'], d['cpu_baseline'])
p=d['pileup']; print({k:p[k] for k in ('value','e2e','e2e_cli','clocks')}); print(p['parity'], p['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ad_ref.json 2> gpurun_out/bench_ad_ref.err; echo "ref arm rc=$?"; cut -c1-200 gpurun_out/bench_ad_ref.json
