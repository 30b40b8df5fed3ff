#!/bin/bash
# round 2, call AH: ncu full capture of k_matesw (mate-rescue local alignment, one lane per SSE element); smoke() on the final build
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ah.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_ah.log | cut -c1-200
timeout 500 ncu --set full --clock-control none -k regex:"k_matesw" --launch-skip 1 -c 1 -f -o /tmp/ms_ah python tools/dpbench.py 100 100000 > gpurun_out/ncu_ms_ah.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ms_ah.log | cut -c1-200
ncu -i /tmp/ms_ah.ncu-rep --page raw --csv > gpurun_out/k_matesw_r02_ah_raw.csv 2>/dev/null; wc -c gpurun_out/k_matesw_r02_ah_raw.csv
