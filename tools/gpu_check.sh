#!/bin/bash
# One-shot GPU verification used at the end of round 1 (run through gpurun): parity tests, the k_seed2 prefetch A/B,
# the pileup bench and the ncu captures of the pileup kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 480 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 240 python tools/ab_seed.py > gpurun_out/ab_seed.log 2> gpurun_out/ab_seed.err; tail -5 gpurun_out/ab_seed.log
timeout 240 python bench.py --path pileup --steps 3 --warmup 3 > gpurun_out/bench_pileup.json 2> gpurun_out/bench_pileup.err; tail -c 600 gpurun_out/bench_pileup.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_plp_win -c 1 -f -o gpurun_out/k_plp_win python bench.py --path pileup --steps 1 --warmup 3 --no-cpu-baseline --no-cli > gpurun_out/ncu_plp.log 2>&1
ncu -i gpurun_out/k_plp_win.ncu-rep --page raw --csv > gpurun_out/k_plp_win_raw.csv 2>/dev/null
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_plp -c 80 --csv --log-file gpurun_out/launches_plp.csv python bench.py --path pileup --steps 2 --warmup 3 --no-cpu-baseline --no-cli > gpurun_out/ncu_plp2.log 2>&1
ls -la gpurun_out
