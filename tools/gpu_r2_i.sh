#!/bin/bash
# round 2, call I: ncu evidence for profiles/ (launch lists + full captures of the dominant kernels)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pileup --no-index-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
KAB="python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_seed2|^k_region" -c 2 -f -o gpurun_out/p_seed_region_r02 $KAB > gpurun_out/ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:k_chain_warp -c 7 -f -o gpurun_out/p_chain_r02 $KAB > gpurun_out/ncu_b.log 2>&1; echo "ncu b rc=$?"
PLP="python bench.py --path pileup --steps 1 --warmup 0 --no-cpu-baseline --no-cli"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_plp_win -s 2 -c 1 -f -o gpurun_out/p_plp_r02 $PLP > gpurun_out/ncu_c.log 2>&1; echo "ncu c rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_plp -c 200 --csv --log-file gpurun_out/launches_plp_r02.csv $PLP > gpurun_out/ncu_d.log 2>&1; echo "ncu d rc=$?"
for f in p_seed_region_r02 p_chain_r02 p_plp_r02; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ncu -i gpurun_out/p_seed_region_r02.ncu-rep --page source --csv > gpurun_out/p_seed_region_r02_source.csv 2>/dev/null
ls -la gpurun_out | tail -15
