#!/bin/bash
# round 2, call H: k_region with next-task prefetch (L2 / L1), and with the packed-reference lines of every seed prefetched too
mkdir -p gpurun_out
timeout 400 python tools/kab.py variants/libbsq_nopf.so,variants/libbsq_pfl2.so,variants/libbsq_pfl1.so,variants/libbsq_pfpac.so 3100 100000 3 > gpurun_out/kab_h.jsonl 2> gpurun_out/kab_h.err; echo "kab rc=$?"
cat gpurun_out/kab_h.jsonl; tail -3 gpurun_out/kab_h.err
