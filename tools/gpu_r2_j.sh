#!/bin/bash
# round 2, call J: per-pass seeding kernels (bsq_seed3.cuh) -- parity tests, then A/B against k_seed2 at bench scale
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_j.log
timeout 900 python tools/kab.py biscuit_b200/csrc/libbsq.so@BSQ_SEED_IMPL=2,biscuit_b200/csrc/libbsq.so 3100 100000 3 > gpurun_out/kab_j.jsonl 2> gpurun_out/kab_j.err; echo "kab rc=$?"; cat gpurun_out/kab_j.jsonl; tail -3 gpurun_out/kab_j.err
