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
    dx.close()
    print(f"smoke ok: {2 * n} tasks, {len(regs)} regions identical to the reference; kernel us seed/sa/chain/extend = {c[5:9]}")
