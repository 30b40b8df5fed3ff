"""__graft_entry__.smoke(): one small phase-1 alignment batch on cuda:0 through the C ABI, checked
against the golden vectors produced by the unmodified reference (tests/golden/align_tiny.npz)."""
import numpy as np

import golden_io
import refprobe
from biscuit_b200 import capi


def run(bsq=None):
    bsq = bsq or capi.load()
    hi, z = golden_io.load_align_tiny()
    dx = bsq.upload(hi, 0)
    opt = bsq.default_opt()
    al = capi.Aligner(dx, opt)
    n = len(z["lens"])
    seqs = np.concatenate([z["seqs"], z["seqs"]])
    lens = np.concatenate([z["lens"], z["lens"]])
    par = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
    regs, off = al.phase1(seqs, lens, par)
    got = refprobe.regs_from_bsq(regs)
    assert (off == z["reg_off"]).all(), "region counts differ from the reference"
    assert (got == z["regs"]).all(), "regions differ from the reference"
    out, n_out = dx.collect_intv(opt, seqs, lens, par)
    io = z["intv_off"]
    for t in range(2 * n):
        if lens[t] >= opt.min_seed_len:
            assert n_out[t] == io[t + 1] - io[t] and (out[t, :n_out[t]] == z["intv"][io[t]:io[t + 1]]).all(), t
    c = al.counters()
    al.close()
    # the batched phase-2 DP: a final CIGAR for every region (k_cigar) and a mate-rescue window per region (k_matesw)
    n_dp = _dp_check(bsq, dx, opt, hi, seqs, lens, regs, off)
    dx.close()
    print(f"smoke ok: {2 * n} tasks, {len(regs)} regions identical to the reference; kernel us seed/sa/chain/extend = {c[5:9]}; "
          f"{n_dp[0]} final CIGARs (k_cigar: every one spans its query and reference interval) and {n_dp[1]} rescue alignments (k_matesw) on the device; "
          "their parity with the reference's bis_bwa_gen_cigar2 / ksw_align2 is tests/test_dp.py")


def _dp_check(bsq, dx, opt, hi, seqs, lens, regs, off):
    task = np.repeat(np.arange(len(lens)), np.diff(off))
    ok = ((regs["rb"] < hi.l_pac) == (regs["re"] <= hi.l_pac)) & (regs["re"] - regs["rb"] <= 1024)
    sel = np.nonzero(ok)[0][:512]
    jobs = np.zeros(len(sel), dtype=capi.CIGAR_JOB_DTYPE)
    r = regs[sel]
    jobs["rb"], jobs["re"], jobs["row"], jobs["qb"], jobs["qe"], jobs["truesc"], jobs["parent"] = r["rb"], r["re"], task[sel], r["qb"], r["qe"], r["truesc"], r["parent"]
    jobs["w"] = np.minimum(np.abs((r["re"] - r["rb"]) - (r["qe"] - r["qb"])) + 3, opt.w << 2)
    dp = capi.Dp(dx, opt)
    dp.set_reads(seqs, lens)
    res, blob = dp.cigar(jobs)
    assert (res["n_cigar"] > 0).all(), "k_cigar returned no alignment for a phase-1 region"
    for k in range(len(sel)):  # every CIGAR spans its query and its reference interval (a trimmed end deletion accounted for)
        w = blob[int(res[k]["off"]):int(res[k]["off"]) + int(res[k]["n_cigar"])]
        ql = int(sum(x >> 4 for x in w if x & 0xf in (0, 1)))
        rl = int(sum(x >> 4 for x in w if x & 0xf in (0, 2)))
        assert ql == jobs[k]["qe"] - jobs[k]["qb"], k
        assert rl <= jobs[k]["re"] - jobs[k]["rb"] and rl + int(res[k]["lead_del"]) <= jobs[k]["re"] - jobs[k]["rb"], k
    mj = np.zeros(len(sel), dtype=capi.MATESW_JOB_DTYPE)
    fwd = r["rb"] < hi.l_pac
    rb = np.where(fwd, 2 * hi.l_pac - r["re"] - 300, 2 * hi.l_pac - r["re"] - 300)
    rb = np.clip(rb, 0, 2 * hi.l_pac - 700)
    rb = np.where((rb < hi.l_pac) != (rb + 600 <= hi.l_pac), 0, rb)
    mj["rb"], mj["re"], mj["row"], mj["use_ga"] = rb, rb + 600, task[sel], r["parent"]
    mj["xtra"] = 0x40000 | 0x80000 | 0x10000 | (opt.min_seed_len * opt.a)
    out = dp.matesw(mj)
    assert (out["pad_"] == 0).all() and (out["score"] >= 0).all()
    dp.close()
    return len(sel), len(mj)
