#!/bin/bash
# round 2, call U: timing of the phase-2 DP kernels at bench scale, ncu full captures (phase-1 kernels, k_cigar, k_matesw),
# launch list of one bench step, and the two-contexts A/B of the end-to-end pipeline
mkdir -p gpurun_out
timeout 600 python tools/dpbench.py 3100 100000 > gpurun_out/dpbench_u.json 2> gpurun_out/dpbench_u.err; echo "dpbench rc=$?"; cat gpurun_out/dpbench_u.json; tail -3 gpurun_out/dpbench_u.err
# full captures of the phase-1 kernels of the second warm run (without the derived full SA: kernel replay cannot save 99 GB)
BSQ_FULL_SA=0 timeout 900 ncu --set full --clock-control none -k regex:"k_s3_|k_seed_sort|k_region|k_chain_warp" --launch-skip 15 -c 15 -f -o /tmp/p1_u python tools/kbench.py biscuit_b200/csrc/libbsq.so > gpurun_out/ncu_p1_u.log 2>&1; echo "ncu p1 rc=$?"; tail -2 gpurun_out/ncu_p1_u.log | cut -c1-200
ncu -i /tmp/p1_u.ncu-rep --page raw --csv > gpurun_out/phase1_kernels_r02_u_raw.csv 2>/dev/null; wc -c gpurun_out/phase1_kernels_r02_u_raw.csv
timeout 600 ncu --set full --clock-control none -k regex:"k_cigar|k_matesw" --launch-skip 2 -c 4 -f -o /tmp/dp_u python tools/dpbench.py 100 100000 > gpurun_out/ncu_dp_u.log 2>&1; echo "ncu dp rc=$?"; tail -2 gpurun_out/ncu_dp_u.log | cut -c1-200
ncu -i /tmp/dp_u.ncu-rep --page raw --csv > gpurun_out/dp_kernels_r02_u_raw.csv 2>/dev/null; wc -c gpurun_out/dp_kernels_r02_u_raw.csv
# launch list of the default bench (align leg)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_(s3|seed_sort|expand|sa|chain|region|compact|cigar|matesw)" -c 4000 --csv --log-file gpurun_out/launches_r02_u.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pileup > gpurun_out/ncu_launch_u.log 2>&1; echo "ncu launches rc=$?"
# two aligner contexts on one device (copies of one batch under the kernels of the other) vs one
for tc in 0 1; do
  if [ $tc = 1 ]; then export BQ_TWO_CONTEXTS=1; fi
  BQ_TIMING=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pileup --no-cpu-baseline > gpurun_out/bench_u_tc$tc.json 2> gpurun_out/bench_u_tc$tc.err; echo "bench tc=$tc rc=$?"
  grep "bq_pipeline" gpurun_out/bench_u_tc$tc.err | tail -1
  python -c "
import json; d=json.load(open('gpurun_out/bench_u_tc$tc.json')); print({k:d[k] for k in ('value','e2e','e2e_phase1','phase2_dp')}); print(d['roofline']['frac'], d['roofline']['moved_layout'], d['roofline']['step'])"
done
