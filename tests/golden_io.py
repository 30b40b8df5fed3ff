"""Load tests/golden/align_tiny.npz (made by tools/make_golden.py from the unmodified reference)."""
import os

import numpy as np

from biscuit_b200.indexio import FmHalf, HostIndex, _aligned_u32

HERE = os.path.dirname(os.path.abspath(__file__))


def load_align_tiny():
    z = np.load(os.path.join(HERE, "golden", "align_tiny.npz"))
    prim0, prim1, sa_intv, seq_len, l_pac = (int(v) for v in z["meta"])
    fm = []
    for w, prim in ((0, prim0), (1, prim1)):
        bwt = _aligned_u32(len(z[f"bwt{w}"]))
        bwt[:] = z[f"bwt{w}"]
        fm.append(FmHalf(prim, z["L2"][w].astype(np.uint64), bwt, z[f"sa{w}"].astype(np.uint64), sa_intv, seq_len))
    hi = HostIndex(fm, z["pac"].copy(), l_pac, [str(s) for s in z["names"]], z["ann_offset"].astype(np.int64),
                   z["ann_len"].astype(np.int32), np.zeros(len(z["ann_len"]), np.int32))
    return hi, z
