#!/bin/bash
# round 2, call AB: where the `biscuit pileup` command line spends its time on the 8 Mb sample (BSQ_PLP_TIMING)
mkdir -p gpurun_out
BSQ_PLP_TIMING=1 timeout 900 python bench.py --path pileup --plp-mb 32 --steps 3 --warmup 3 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
grep "^\[pileup\]\|Real time" gpurun_out/bench_ab.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); print({k:d.get(k) for k in ('value','e2e','e2e_cli','cpu_baseline')})"
