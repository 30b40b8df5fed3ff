#!/bin/bash
# round 2, call E: persistent warps (chain tiers, region), seed ring of 8, two-count extension, one-block-at-a-time variants
mkdir -p gpurun_out
timeout 600 python tools/kab.py variants/libbsq_base.so,variants/libbsq_p1.so,variants/libbsq_p2.so,variants/libbsq_p3.so,variants/libbsq_p4.so,variants/libbsq_p5.so > gpurun_out/kab_e.jsonl 2> gpurun_out/kab_e.err
cat gpurun_out/kab_e.jsonl
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_align_sam.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e.log
tail -5 gpurun_out/pytest_e.log
