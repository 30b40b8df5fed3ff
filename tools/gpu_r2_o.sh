#!/bin/bash
# round 2, call O: k_region with per-chain inputs prepared one chain per lane (k_region_prep); occupancy variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_o.log
timeout 900 python tools/kab.py biscuit_b200/csrc/libbsq.so,variants/libbsq_rc7.so,variants/libbsq_rc6.so 3100 100000 3 > gpurun_out/kab_o.jsonl 2> gpurun_out/kab_o.err; echo "kab rc=$?"; cat gpurun_out/kab_o.jsonl; tail -3 gpurun_out/kab_o.err
KAB="python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_region" -c 4 --csv --log-file gpurun_out/launches_region_o.csv $KAB > gpurun_out/ncu_o0.log 2>&1; echo "launch list rc=$?"
grep -o '"k_region[a-z_]*\|ns","[0-9.]*' gpurun_out/launches_region_o.csv | paste - - | head
