#!/bin/bash
# round 2, call V: align tests + align bench after the deferred fetch / parallel pestat / parallel preparation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_sam.py tests/test_dp.py tests/test_edges.py tests/test_boundary.py tests/test_phase1.py -m gpu -q -x > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v.log
tail -4 gpurun_out/pytest_v.log | cut -c1-300
BQ_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pileup > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
grep "bq_pipeline\|parity_at_scale" gpurun_out/bench_v.err | tail -3 | cut -c1-300
grep "bq_finish_a\|bq_finish_b\|bq_batch_run" gpurun_out/bench_v.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print({k:d[k] for k in ('value','e2e','e2e_phase1','phase2_dp')}); print(d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'])"
