#!/bin/bash
# round 2, call K: where the per-pass seeding kernels spend their time
mkdir -p gpurun_out
KAB="python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_s3|^k_seed_sort" -c 40 --csv --log-file gpurun_out/launches_seed3_k.csv $KAB > gpurun_out/ncu_k0.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_s3" -c 5 -f -o gpurun_out/p_seed3_k $KAB > gpurun_out/ncu_k1.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/p_seed3_k.ncu-rep --page raw --csv > gpurun_out/p_seed3_k_raw.csv 2>/dev/null
ncu -i gpurun_out/p_seed3_k.ncu-rep --page source --csv > gpurun_out/p_seed3_k_source.csv 2>/dev/null
rm -f gpurun_out/p_seed3_k.ncu-rep
ls -la gpurun_out | tail
