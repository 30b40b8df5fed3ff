#!/bin/bash
# round 2, call S: the batched phase-2 DP kernels (k_cigar, k_matesw) against the reference's bis_bwa_gen_cigar2 / ksw_align2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp.py -m gpu -x -q > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_s.log | cut -c1-400
