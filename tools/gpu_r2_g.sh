#!/bin/bash
# round 2, call G: k_region_grp (four tasks per warp, DP rows convergent) against the warp-per-task kernel
mkdir -p gpurun_out
timeout 300 python tools/kab.py variants/libbsq_cur.so,variants/libbsq_g1.so 3100 100000 3 > gpurun_out/kab_g.jsonl 2> gpurun_out/kab_g.err; echo "kab rc=$?"
cat gpurun_out/kab_g.jsonl; tail -3 gpurun_out/kab_g.err
