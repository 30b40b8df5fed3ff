#!/bin/bash
# round 2, call AG: pileup tests and command-line timing with the slab allocator for the page-locked batch arrays
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pileup_cli.py tests/test_pileup_ref.py tests/test_pileup.py -m gpu -q -x > gpurun_out/pytest_ag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ag.log | cut -c1-200
BSQ_PLP_TIMING=1 timeout 600 python bench.py --path pileup --plp-mb 32 --steps 3 --warmup 3 > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err; echo "bench rc=$?"
grep "^\[pileup\]\|Real time" gpurun_out/bench_ag.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_ag.json')); print({k:d.get(k) for k in ('e2e_cli','parity')})"
