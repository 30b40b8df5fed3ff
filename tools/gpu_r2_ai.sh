#!/bin/bash
# round 2, call AI: the result-slot test and the multi-batch SAM tests on the CUDA build
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_phase1.py tests/test_align_sam.py -m gpu -q -x -k "deferred or many_small or dp_is_used or smart" > gpurun_out/pytest_ai.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ai.log | cut -c1-300
