#!/bin/bash
# round 2, call N: source-level profiles of k_region and the chain tiers (instruction counts per CUDA line)
mkdir -p gpurun_out
KAB="python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_region|k_chain_warp" -c 8 -f -o gpurun_out/p_rc_n $KAB > gpurun_out/ncu_n1.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/p_rc_n.ncu-rep --page raw --csv > gpurun_out/p_rc_n_raw.csv 2>/dev/null
ncu -i gpurun_out/p_rc_n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/p_rc_n_source.csv 2>gpurun_out/ncu_n2.log || ncu -i gpurun_out/p_rc_n.ncu-rep --page source --csv > gpurun_out/p_rc_n_source.csv 2>>gpurun_out/ncu_n2.log
rm -f gpurun_out/p_rc_n.ncu-rep
timeout 300 python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 3 > gpurun_out/kab_n.jsonl 2>/dev/null; cat gpurun_out/kab_n.jsonl
ls -la gpurun_out | tail -8
