#!/bin/bash
# round 2, call AL: bench.py after the byte-accounting edit (align leg)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-pileup > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_al.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_al.json')); print(d['value'], d['e2e'], d['parity_at_scale']['identical'])"
