"""ctypes access to oracle/_ref/libbiscuit_ref.so (the UNMODIFIED reference compiled by
oracle/Makefile) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbiscuit_ref.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "biscuit_ref")


def available() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefProbe:
    def __init__(self, prefix: str):
        self.lib = C.CDLL(REF_SO)
        self.lib.refp_open.restype = C.c_void_p
        self.lib.refp_open.argtypes = [C.c_char_p]
        self.lib.refp_chain.restype = C.c_int64
        self.h = C.c_void_p(self.lib.refp_open(prefix.encode()))
        if not self.h:
            raise RuntimeError("refp_open failed")

    def close(self):
        self.lib.refp_close(self.h)

    def info(self):
        out = np.zeros(15, dtype=np.int64)
        self.lib.refp_index_info(self.h, _p(out))
        return out

    def occ4(self, which, k):
        k = np.ascontiguousarray(k, dtype=np.int64)
        out = np.zeros((len(k), 4), dtype=np.uint64)
        self.lib.refp_occ4(self.h, C.c_int(which), C.c_int(len(k)), _p(k), _p(out))
        return out

    def sa(self, which, k):
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.zeros(len(k), dtype=np.uint64)
        self.lib.refp_sa(self.h, C.c_int(which), C.c_int(len(k)), _p(k), _p(out))
        return out

    def collect_intv(self, parent, seq, cap=4096):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.zeros((cap, 4), dtype=np.uint64)
        n = self.lib.refp_collect_intv(self.h, C.c_int(parent), C.c_int(len(seq)), _p(seq), _p(out), C.c_int(cap))
        return out[:n].copy()

    def chain(self, parent, seq, stage=1, cap=1 << 20):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.int64)
        n_ch = C.c_int()
        fr = C.c_float()
        o = self.lib.refp_chain(self.h, C.c_int(parent), C.c_int(len(seq)), _p(seq), C.c_int(stage), C.byref(n_ch),
                                C.byref(fr), _p(out), C.c_int64(cap))
        assert o >= 0
        return out[:o].copy(), n_ch.value, fr.value

    def align1(self, parent, seq, cap=4096):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.zeros((cap, 16), dtype=np.int64)
        n = self.lib.refp_align1(self.h, C.c_int(parent), C.c_int(len(seq)), _p(seq), _p(out), C.c_int(cap))
        return out[:n].copy()

    def worker1(self, which_read, seq, cap=4096):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.zeros((cap, 16), dtype=np.int64)
        n = self.lib.refp_worker1(self.h, C.c_int(which_read), C.c_int(len(seq)), _p(seq), _p(out), C.c_int(cap))
        return out[:n].copy()

    def extend2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        out = np.zeros(6, dtype=np.int32)
        self.lib.refp_extend2(C.c_int(len(q)), _p(q), C.c_int(len(t)), _p(t), _p(mat), C.c_int(o_del), C.c_int(e_del),
                              C.c_int(o_ins), C.c_int(e_ins), C.c_int(w), C.c_int(end_bonus), C.c_int(zdrop), C.c_int(h0),
                              _p(out))
        return out


def regs_from_ref(rows: np.ndarray) -> np.ndarray:
    """refp 16-column rows -> comparable tuple array (rb,re,qb,qe,rid,score,truesc,w,seedcov,seedlen0,frac_bits,bss,parent)."""
    out = np.zeros((len(rows), 13), dtype=np.int64)
    if len(rows):
        out[:, 0:7] = rows[:, 0:7]
        out[:, 7] = rows[:, 10]
        out[:, 8] = rows[:, 11]
        out[:, 9] = rows[:, 12]
        out[:, 10] = rows[:, 15]
        out[:, 11] = rows[:, 14] & 1
        out[:, 12] = rows[:, 14] >> 1
    return out


def regs_from_bsq(regs: np.ndarray) -> np.ndarray:
    out = np.zeros((len(regs), 13), dtype=np.int64)
    for i, f in enumerate(("rb", "re", "qb", "qe", "rid", "score", "truesc", "w", "seedcov", "seedlen0")):
        out[:, i] = regs[f]
    out[:, 10] = regs["frac_rep"].view(np.uint32)
    out[:, 11] = regs["bss"]
    out[:, 12] = regs["parent"]
    return out
