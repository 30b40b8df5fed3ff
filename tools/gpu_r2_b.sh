#!/bin/bash
# round 2, call B: the >2^32 index test and the new bench line (index check, parity at scale, pileup leg)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_index_build.py -m gpu -q -x -k beyond --durations=5 > gpurun_out/pytest_index.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_index.log
tail -8 gpurun_out/pytest_index.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
grep -v "mem_pestat\|^\[M::" gpurun_out/bench_b.err | tail -30
tail -c 3000 gpurun_out/bench_b.json
