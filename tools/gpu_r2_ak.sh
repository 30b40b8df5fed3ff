#!/bin/bash
# round 2, call AK: last check of the command-line tests on the CUDA build after the host-side changes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_align_sam.py tests/test_pileup_cli.py -m gpu -q -x > gpurun_out/pytest_ak.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ak.log | cut -c1-200
