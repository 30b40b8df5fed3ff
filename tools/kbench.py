"""Kernel-only timing of phase 1 with an alternative build of libbsq (tuning experiments):
    python tools/kbench.py <lib.so> [ref_mb] [pairs]
prints the per-kernel device times of one warm run."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from biscuit_b200 import capi  # noqa: E402

libs = sys.argv[1].split(",")
ref_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 3100
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
nt4, pac, names, offs, lens = bench.gen_reference(ref_mb)
reads = bench.sim_batch(nt4, names, offs, lens, pairs, seed=2024)
seqs, tl, par = bench.tasks_from_reads(reads)
for lib in libs:
    bsq = capi.Bsq(lib)
    dx = bsq.build_index(pac, len(nt4), names, offs, lens, device=0)
    al = capi.Aligner(dx, bsq.default_opt())
    for it in range(3):
        al.phase1(seqs, tl, par)
        c = al.counters()
    print(os.path.basename(lib), dict(zip(["k_seed", "k_sa", "k_chain", "k_region", "scan", "all"], [int(x) for x in c[5:11]])), flush=True)
    al.close()
    dx.close()
