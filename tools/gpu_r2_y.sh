#!/bin/bash
# round 2, call Y: align tests + align bench after the SAM-text and reader changes; CLI throughput on a 2 M-pair FASTQ
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_sam.py tests/test_edges.py tests/test_boundary.py -m gpu -q -x > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_y.log
tail -3 gpurun_out/pytest_y.log | cut -c1-300
BQ_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pileup > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "bench rc=$?"
grep "bq_pipeline\|parity_at_scale" gpurun_out/bench_y.err | tail -2 | cut -c1-300
grep "bq_finish_a\|bq_finish_b" gpurun_out/bench_y.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/bench_y.json')); print({k:d[k] for k in ('value','e2e','e2e_phase1')})"
