"""A/B of the k_seed2 variants (BSQ_SEED_VARIANT=1/0, list in AB_MODES) inside one process, on the default bench workload:
    python tools/ab_seed.py [ref_mb] [pairs]
prints the per-kernel device times of warm runs for both settings and checks that the regions are identical."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from biscuit_b200 import capi  # noqa: E402

ref_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 3100
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
nt4, pac, names, offs, lens = bench.gen_reference(ref_mb)
reads = bench.sim_batch(nt4, names, offs, lens, pairs, seed=2024)
seqs, tl, par = bench.tasks_from_reads(reads)
bsq = capi.load()
dx = bsq.build_index(pac, len(nt4), names, offs, lens, device=0)
al = capi.Aligner(dx, bsq.default_opt())
out, ref = [], None
for pf in [int(x) for x in os.environ.get("AB_MODES", "1,0,1,0").split(",")]:
    os.environ["BSQ_SEED_VARIANT"] = str(pf)
    for it in range(3):
        regs, off = al.phase1(seqs, tl, par)
        c = al.counters()
    same = None
    if ref is None:
        ref = (regs.tobytes(), off.tobytes())
    else:
        same = ref == (regs.tobytes(), off.tobytes())
    row = {"variant": pf, "identical_to_first": same, **dict(zip(["k_seed", "k_sa", "k_chain", "k_region", "scan", "all"], [int(x) for x in c[5:11]]))}
    out.append(row)
    print(json.dumps(row), flush=True)
al.close()
dx.close()
