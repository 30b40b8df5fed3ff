#!/bin/bash
# round 2, call Q: chain filter order -- Hoare partitions of ks_introsort done by the whole warp vs replayed on lane 0
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
timeout 900 python tools/kab.py variants/libbsq_seqpart.so,biscuit_b200/csrc/libbsq.so 3100 100000 3 > gpurun_out/kab_q.jsonl 2> gpurun_out/kab_q.err; echo "kab rc=$?"; cat gpurun_out/kab_q.jsonl; tail -3 gpurun_out/kab_q.err
