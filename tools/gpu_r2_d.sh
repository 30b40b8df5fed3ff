#!/bin/bash
# round 2, call D: A/B of region v2 (lane-parallel control), chain filter sweeps + register sort, seed shared-memory sizes
mkdir -p gpurun_out
timeout 600 python tools/kab.py variants/libbsq_base.so,variants/libbsq_r4.so,variants/libbsq_c2.so,variants/libbsq_c3.so,variants/libbsq_s45.so,variants/libbsq_s85l.so > gpurun_out/kab_d.jsonl 2> gpurun_out/kab_d.err
cat gpurun_out/kab_d.jsonl
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_align_sam.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -5 gpurun_out/pytest_d.log
