"""A/B of alternative builds of libbsq.so on the bench workload, in one process:
    python tools/kab.py <lib1.so,lib2.so,...> [ref_mb] [pairs] [reps]
Per library: per-kernel device times of warm runs (min and median over `reps`) and a digest of the regions; the
first library is the reference for "identical".  Used to try kernel variants built with different -D flags
(tools/build_variants.sh) in a single gpurun call."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from biscuit_b200 import capi  # noqa: E402

libs = sys.argv[1].split(",")
ref_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 3100
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
nt4, pac, names, offs, lens = bench.gen_reference(ref_mb)
reads = bench.sim_batch(nt4, names, offs, lens, pairs, seed=2024)
seqs, tl, par = bench.tasks_from_reads(reads)
first = None
for lib in libs:
    # "lib.so@NAME=VALUE[@NAME=VALUE]" sets environment switches the library reads at launch time (e.g. BSQ_SEED_IMPL=2)
    lib, *envs = lib.split("@")
    for kv in envs:
        os.environ[kv.split("=")[0]] = kv.split("=")[1]
    path = lib if os.path.isabs(lib) else os.path.join(ROOT, lib)
    try:
        bsq = capi.Bsq(path)
        dx = bsq.build_index(pac, len(nt4), names, offs, lens, device=0)
        al = capi.Aligner(dx, bsq.default_opt())
        rows = []
        for it in range(reps + 1):
            regs, off = al.phase1(seqs, tl, par)
            if it:
                rows.append(al.counters()[5:11].astype(np.int64))
        dig = hashlib.sha1(regs.tobytes() + off.tobytes()).hexdigest()[:16]
        if first is None:
            first = dig
        rows = np.array(rows)
        names_k = ["k_seed", "k_sa", "k_chain", "k_region", "scan", "all"]
        print(json.dumps({"lib": os.path.basename(lib) + "".join("@" + e for e in envs), "identical": dig == first, "n_regs": int(len(regs)),
                          "min_us": dict(zip(names_k, rows.min(axis=0).tolist())), "median_us": dict(zip(names_k, np.median(rows, axis=0).astype(int).tolist()))}),
              flush=True)
        al.close()
        dx.close()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"lib": os.path.basename(lib), "error": str(e)}), flush=True)
    for kv in envs:
        os.environ.pop(kv.split("=")[0], None)
