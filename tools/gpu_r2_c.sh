#!/bin/bash
# round 2, call C: A/B of kernel variants (region v2, seed occupancy, LDG256) + parity tests of the new region kernel
mkdir -p gpurun_out
timeout 600 python tools/kab.py variants/libbsq_base.so,variants/libbsq_r2.so,variants/libbsq_r2c6.so,variants/libbsq_ldg256.so,variants/libbsq_s165.so,variants/libbsq_s86.so,variants/libbsq_s87.so > gpurun_out/kab_c.jsonl 2> gpurun_out/kab_c.err
cat gpurun_out/kab_c.jsonl
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_align_sam.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
