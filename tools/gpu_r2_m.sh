#!/bin/bash
# round 2, call M: random-gather ceiling of the device + occupancy sweep of the seeding kernels
mkdir -p gpurun_out
timeout 600 tools/micro/gather_peak > gpurun_out/gather_peak_m.jsonl 2>&1; echo "gather rc=$?"; cat gpurun_out/gather_peak_m.jsonl
timeout 900 python tools/kab.py variants/libbsq_s3c6.so,variants/libbsq_s3c5.so,variants/libbsq_s3c4.so,variants/libbsq_s3c3.so 3100 100000 3 > gpurun_out/kab_m.jsonl 2> gpurun_out/kab_m.err; echo "kab rc=$?"; cat gpurun_out/kab_m.jsonl; tail -3 gpurun_out/kab_m.err
