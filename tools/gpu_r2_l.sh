#!/bin/bash
# round 2, call L: seeding kernels with prefetched bases / register-held column heads / deferred stores; occupancy variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase1.py tests/test_edges.py -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_l.log
timeout 900 python tools/kab.py biscuit_b200/csrc/libbsq.so@BSQ_SEED_IMPL=2,biscuit_b200/csrc/libbsq.so,variants/libbsq_s3c7.so,variants/libbsq_s3c6.so 3100 100000 3 > gpurun_out/kab_l.jsonl 2> gpurun_out/kab_l.err; echo "kab rc=$?"; cat gpurun_out/kab_l.jsonl; tail -3 gpurun_out/kab_l.err
KAB="python tools/kab.py biscuit_b200/csrc/libbsq.so 3100 100000 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_s3|^k_seed_sort" -c 12 --csv --log-file gpurun_out/launches_seed3_l.csv $KAB > gpurun_out/ncu_l0.log 2>&1; echo "launch list rc=$?"
grep -o '"k_s3[^"]*\|"void k_s3[^"]*\|"k_seed_sort[^"]*\|"[0-9]*","ns\|ns","[0-9.]*' gpurun_out/launches_seed3_l.csv | paste - - | head -12
