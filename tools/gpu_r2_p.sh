#!/bin/bash
# round 2, call P: full default bench (align + pileup legs) with the new seeding / region kernels
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; tail -25 gpurun_out/bench_p.err | cut -c1-400; cat gpurun_out/bench_p.json | cut -c1-6000
