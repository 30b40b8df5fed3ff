"""Kernel-only timing of the batched phase-2 DP (k_cigar, k_matesw) at the bench's scale:
    python tools/dpbench.py [ref_mb] [pairs]
Phase 1 of one synthetic batch gives the regions; job set `best` = the best region of every task (what phase 2 asks for on
clean pairs: mostly ungapped), `all` = every region with score >= 30 (partial hits: banded global alignment + traceback),
`matesw` = one mate-rescue window (600 bases around the hit, opposite strand) per read with a hit.  Prints one JSON line;
`ncu -k regex:k_cigar|k_matesw` over this script gives the captures under profiles/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from biscuit_b200 import capi  # noqa: E402


def infer_bw(l1, l2, score, a, q, r):
    w = np.floor((np.minimum(l1, l2) * a - score - q) / r + 2.).astype(np.int64)
    w = np.maximum(w, np.abs(l1 - l2))
    return np.where((l1 == l2) & (l1 * a - score < (q + r - a) << 1), 0, w)


def main():
    ref_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 3100
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    nt4, pac, names, offs, lens = bench.gen_reference(ref_mb)
    reads = bench.sim_batch(nt4, names, offs, lens, pairs, seed=2024)
    seqs, tl, par = bench.tasks_from_reads(reads)
    bsq = capi.load()
    dx = bsq.build_index(pac, len(nt4), names, offs, lens, device=0)
    opt = bsq.default_opt()
    al = capi.Aligner(dx, opt)
    regs, off = al.phase1(seqs, tl, par)
    al.close()
    l_pac = len(nt4)
    task = np.repeat(np.arange(len(tl)), np.diff(off))
    l1 = (regs["qe"] - regs["qb"]).astype(np.int64)
    l2 = (regs["re"] - regs["rb"]).astype(np.int64)
    w = np.maximum(infer_bw(l1, l2, regs["truesc"].astype(np.int64), opt.a, opt.o_del, opt.e_del),
                   infer_bw(l1, l2, regs["truesc"].astype(np.int64), opt.a, opt.o_ins, opt.e_ins))
    w = np.where(w > opt.w, np.minimum(w, regs["w"]), w)
    rev = regs["rb"] >= l_pac
    jobs = np.zeros(len(regs), dtype=capi.CIGAR_JOB_DTYPE)
    jobs["rb"], jobs["re"], jobs["row"], jobs["qb"], jobs["qe"], jobs["w"], jobs["truesc"] = regs["rb"], regs["re"], task, regs["qb"], regs["qe"], w, regs["truesc"]
    jobs["clip5"] = np.where(rev, tl[task] - regs["qe"], regs["qb"])
    jobs["clip3"] = np.where(rev, regs["qb"], tl[task] - regs["qe"])
    jobs["parent"] = regs["parent"]
    ok = (l2 <= 1024) & ((regs["rb"] < l_pac) == (regs["re"] <= l_pac))
    # best region per task
    order = np.lexsort((-regs["score"], task))
    first = np.ones(len(order), bool)
    first[1:] = task[order][1:] != task[order][:-1]
    best = order[first]
    best = best[ok[best] & (regs["score"][best] >= 30)]
    allj = np.nonzero(ok & (regs["score"] >= 30))[0]
    dp = capi.Dp(dx, opt)
    dp.set_reads(seqs, tl)
    out = {"ref_mb": ref_mb, "tasks": int(len(tl)), "regions": int(len(regs))}
    for name, sel in (("best", best), ("all", allj)):
        ms = []
        for _ in range(3):
            res, blob = dp.cigar(jobs[sel])
            c = dp.counters()
            ms.append(c[3] / 1000)
        gapped = int((res["n_cigar"] > 1 + (jobs[sel]["clip5"] > 0) + (jobs[sel]["clip3"] > 0)).sum())
        out["cigar_" + name] = {"jobs": int(len(sel)), "ungapped_first_try": int(c[1]), "dp_cells": int(c[2]), "kernel_ms": min(ms), "gapped_cigars": gapped,
                                "unsupported": int((res["n_cigar"] < 0).sum()), "blob_mb": blob.nbytes / 1e6, "jobs_per_s": len(sel) / (min(ms) * 1e-3)}
    # mate rescue: the mate of every best hit, searched on the opposite strand in a 600-base window next to the hit
    b = best[:: 2]
    rb = np.where(regs["rb"][b] < l_pac, 2 * l_pac - regs["re"][b] - 450, 2 * l_pac - regs["re"][b] - 450)
    rb = np.clip(rb, 0, 2 * l_pac - 700)
    re = rb + 600
    keep = (rb < l_pac) == (re <= l_pac)
    mj = np.zeros(int(keep.sum()), dtype=capi.MATESW_JOB_DTYPE)
    mj["rb"], mj["re"] = rb[keep], re[keep]
    mj["row"] = task[b][keep] ^ 1  # the neighbouring task row (the other conversion / the mate): any read serves as a query
    mj["xtra"] = 0x40000 | 0x80000 | 0x10000 | (opt.min_seed_len * opt.a)
    mj["use_ga"] = regs["parent"][b][keep]
    ms = []
    for _ in range(3):
        r = dp.matesw(mj)
        ms.append(dp.counters()[5] / 1000)
    cells = float((600 * tl[mj["row"]]).sum())
    out["matesw"] = {"jobs": int(len(mj)), "kernel_ms": min(ms), "cells_first_pass": cells, "gcups_first_pass": cells / (min(ms) * 1e-3) / 1e9,
                     "hits": int(((r["score"] >= opt.min_seed_len) & (r["qb"] >= 0)).sum())}
    dp.close()
    dx.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
