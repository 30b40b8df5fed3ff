"""ctypes binding of include/bsq.h.

`load()` binds the product library (biscuit_b200/csrc/libbsq.so, CUDA, sm_100a) and raises if it
is missing -- there is no fallback.  Tests may bind another shared object that exports the same
ABI (tests/hostemu) through `Bsq(path)`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .indexio import HostIndex

MAX_READ_LEN = 256
MAX_INTV = 384
HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libbsq.so")


class BsqError(RuntimeError):
    pass


class Intv(C.Structure):
    _fields_ = [("x", C.c_uint64 * 3), ("info", C.c_uint64)]


class Reg(C.Structure):
    _fields_ = [("rb", C.c_int64), ("re", C.c_int64), ("qb", C.c_int32), ("qe", C.c_int32), ("rid", C.c_int32),
                ("score", C.c_int32), ("truesc", C.c_int32), ("w", C.c_int32), ("seedcov", C.c_int32),
                ("seedlen0", C.c_int32), ("frac_rep", C.c_float), ("bss", C.c_uint8), ("parent", C.c_uint8),
                ("pad_", C.c_uint8 * 2)]


REG_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
                      ("truesc", "<i4"), ("w", "<i4"), ("seedcov", "<i4"), ("seedlen0", "<i4"), ("frac_rep", "<f4"),
                      ("bss", "u1"), ("parent", "u1"), ("pad_", "u1", (2,))])
assert REG_DTYPE.itemsize == C.sizeof(Reg) == 56


class Opt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_clip5", "pen_clip3", "w", "zdrop", "min_seed_len",
                 "split_width", "max_occ", "max_chain_gap", "min_chain_weight", "max_chain_extend", "max_mem_intv",
                 "split_len", "self_ovlp", "bsstrand")] + [("mask_level", C.c_float), ("drop_ratio", C.c_float),
                                                            ("ctmat", C.c_int8 * 25), ("gamat", C.c_int8 * 25),
                                                            ("pad_", C.c_int8 * 2)]


class IndexDesc(C.Structure):
    _fields_ = [("bwt", C.c_void_p * 2), ("bwt_words", C.c_uint64 * 2), ("primary", C.c_uint64 * 2),
                ("L2", (C.c_uint64 * 5) * 2), ("seq_len", C.c_uint64), ("sa", C.c_void_p * 2), ("n_sa", C.c_uint64 * 2),
                ("sa_intv", C.c_int32 * 2), ("pac", C.c_void_p), ("l_pac", C.c_int64), ("n_seqs", C.c_int32),
                ("ann_offset", C.c_void_p), ("ann_len", C.c_void_p), ("ann_is_alt", C.c_void_p)]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Bsq:
    """One loaded implementation of the bsq.h ABI."""

    SYMBOLS = ["bsq_opt_default", "bsq_strerror", "bsq_last_error", "bsq_index_upload", "bsq_index_free", "bsq_occ4",
               "bsq_sa_lookup", "bsq_collect_intv", "bsq_extend_batch", "bsq_aligner_create", "bsq_aligner_destroy",
               "bsq_align_phase1", "bsq_free", "bsq_aligner_counters", "bsq_dp_create", "bsq_dp_destroy", "bsq_dp_set_reads",
               "bsq_dp_sync", "bsq_dp_cigar_submit", "bsq_dp_cigar_wait", "bsq_dp_matesw_submit", "bsq_dp_matesw_wait", "bsq_dp_counters"]

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise BsqError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.bsq_strerror.restype = C.c_char_p
        L.bsq_last_error.restype = C.c_char_p
        for name in self.SYMBOLS:
            getattr(L, name)  # AttributeError if the ABI is incomplete
        L.bsq_free.argtypes = [C.c_void_p]
        L.bsq_index_free.argtypes = [C.c_void_p]
        L.bsq_aligner_destroy.argtypes = [C.c_void_p]
        L.bsq_dp_destroy.argtypes = [C.c_void_p]
        for f in ("bsq_dp_sync", "bsq_dp_matesw_wait"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.bsq_dp_cigar_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

    def build_index(self, pac: np.ndarray, l_pac: int, names, ann_offset, ann_len, ann_is_alt=None, device: int = 0) -> "DevIndex":
        """GPU `biscuit index`: both FM-indices from the forward 2-bit pac; the result stays on the device."""
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        ann_offset = np.ascontiguousarray(ann_offset, dtype=np.int64)
        ann_len = np.ascontiguousarray(ann_len, dtype=np.int32)
        alt = np.zeros(len(ann_len), np.int32) if ann_is_alt is None else np.ascontiguousarray(ann_is_alt, dtype=np.int32)
        out = C.c_void_p()
        rc = self.lib.bsq_index_build(_p(pac), C.c_int64(l_pac), C.c_int32(len(ann_len)), _p(ann_offset), _p(ann_len), _p(alt),
                                      C.c_int(device), C.byref(out))
        self.check(rc, "bsq_index_build")
        return DevIndex(self, out, None)

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise BsqError(f"{what}: {self.lib.bsq_strerror(rc).decode()} ({self.lib.bsq_last_error().decode()})")

    def default_opt(self) -> Opt:
        o = Opt()
        self.lib.bsq_opt_default(C.byref(o))
        return o

    # ---- index ----
    def upload(self, hi: HostIndex, device: int = 0) -> "DevIndex":
        d = IndexDesc()
        for w in (0, 1):
            h = hi.fm[w]
            d.bwt[w] = h.bwt.ctypes.data
            d.bwt_words[w] = len(h.bwt)
            d.primary[w] = h.primary
            for i in range(5):
                d.L2[w][i] = int(h.L2[i])
            d.sa[w] = h.sa.ctypes.data
            d.n_sa[w] = len(h.sa)
            d.sa_intv[w] = h.sa_intv
        d.seq_len = hi.fm[0].seq_len
        d.pac = hi.pac.ctypes.data
        d.l_pac = hi.l_pac
        d.n_seqs = len(hi.names)
        d.ann_offset = hi.ann_offset.ctypes.data
        d.ann_len = hi.ann_len.ctypes.data
        d.ann_is_alt = hi.ann_is_alt.ctypes.data
        out = C.c_void_p()
        self.check(self.lib.bsq_index_upload(C.byref(d), C.c_int(device), C.byref(out)), "bsq_index_upload")
        return DevIndex(self, out, hi)


class DevIndex:
    def __init__(self, bsq: Bsq, handle, host: HostIndex):
        self.bsq, self.h, self.host = bsq, handle, host

    def close(self):
        if self.h:
            self.bsq.lib.bsq_index_free(self.h)
            self.h = None

    def sizes(self):
        bw = np.zeros(2, np.uint64); ns = np.zeros(2, np.uint64); pr = np.zeros(2, np.uint64)
        L2 = np.zeros(10, np.uint64); st = np.zeros(3, np.int64)
        self.bsq.check(self.bsq.lib.bsq_index_sizes(self.h, _p(bw), _p(ns), _p(pr), _p(L2), _p(st)), "bsq_index_sizes")
        return dict(bwt_words=bw, n_sa=ns, primary=pr, L2=L2.reshape(2, 5), stats=st)

    def download(self, which: int):
        """(bwt u32 words, sa u64) of one half, exactly as the reference stores them on disk."""
        sz = self.sizes()
        bwt = np.zeros(int(sz["bwt_words"][which]), np.uint32)
        sa = np.zeros(int(sz["n_sa"][which]), np.uint64)
        self.bsq.check(self.bsq.lib.bsq_index_download(self.h, C.c_int(which), _p(bwt), _p(sa)), "bsq_index_download")
        return bwt, sa

    def occ4(self, which: int, k: np.ndarray) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty((len(k), 4), dtype=np.uint64)
        self.bsq.check(self.bsq.lib.bsq_occ4(self.h, C.c_int(which), C.c_int64(len(k)), _p(k), _p(out)), "bsq_occ4")
        return out

    def sa_lookup(self, which: int, k: np.ndarray) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty(len(k), dtype=np.uint64)
        self.bsq.check(self.bsq.lib.bsq_sa_lookup(self.h, C.c_int(which), C.c_int64(len(k)), _p(k), _p(out)),
                       "bsq_sa_lookup")
        return out

    def collect_intv(self, opt: Opt, seqs: np.ndarray, lens: np.ndarray, parent: np.ndarray):
        """seqs: (n_tasks, stride) uint8 nt4, unconverted.  Returns (intv (n,MAX_INTV,4) u64, n_out)."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        n, stride = seqs.shape
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        parent = np.ascontiguousarray(parent, dtype=np.uint8)
        out = np.zeros((n, MAX_INTV, 4), dtype=np.uint64)
        n_out = np.zeros(n, dtype=np.int32)
        rc = self.bsq.lib.bsq_collect_intv(self.h, C.byref(opt), C.c_int64(n), _p(seqs), C.c_int32(stride), _p(lens),
                                           _p(parent), _p(out), _p(n_out))
        self.bsq.check(rc, "bsq_collect_intv")
        return out, n_out


def extend_batch(bsq: Bsq, opt: Opt, queries, targets, is_parent, w, h0) -> np.ndarray:
    """queries/targets: lists of nt4 uint8 arrays.  Returns (n,6) int32 {score,qle,tle,gtle,gscore,max_off}."""
    n = len(queries)
    qlen = np.array([len(q) for q in queries], dtype=np.int32)
    tlen = np.array([len(t) for t in targets], dtype=np.int32)
    qoff = np.concatenate([[0], np.cumsum(qlen)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
    toff = np.concatenate([[0], np.cumsum(tlen)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
    qbuf = np.concatenate(queries).astype(np.uint8) if n else np.zeros(1, np.uint8)
    tbuf = np.concatenate(targets).astype(np.uint8) if n else np.zeros(1, np.uint8)
    if len(qbuf) == 0:
        qbuf = np.zeros(1, np.uint8)
    if len(tbuf) == 0:
        tbuf = np.zeros(1, np.uint8)
    is_parent = np.ascontiguousarray(is_parent, dtype=np.uint8)
    w = np.ascontiguousarray(w, dtype=np.int32)
    h0 = np.ascontiguousarray(h0, dtype=np.int32)
    out = np.zeros((n, 6), dtype=np.int32)
    rc = bsq.lib.bsq_extend_batch(C.byref(opt), C.c_int64(n), _p(qbuf), _p(qoff), _p(qlen), _p(tbuf), _p(toff), _p(tlen),
                                  _p(is_parent), _p(w), _p(h0), _p(out))
    bsq.check(rc, "bsq_extend_batch")
    return out


class Aligner:
    def __init__(self, idx: DevIndex, opt: Opt):
        self.idx, self.bsq, self.opt = idx, idx.bsq, opt
        h = C.c_void_p()
        self.bsq.check(self.bsq.lib.bsq_aligner_create(idx.h, C.byref(opt), C.byref(h)), "bsq_aligner_create")
        self.h = h

    def close(self):
        if self.h:
            self.bsq.lib.bsq_aligner_destroy(self.h)
            self.h = None

    def phase1(self, seqs: np.ndarray, lens: np.ndarray, parent: np.ndarray):
        """Returns (regs structured array, reg_off int64[n_tasks+1])."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        n, stride = seqs.shape
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        parent = np.ascontiguousarray(parent, dtype=np.uint8)
        reg_off = np.zeros(n + 1, dtype=np.int64)
        regs_p = C.c_void_p()
        rc = self.bsq.lib.bsq_align_phase1(self.h, C.c_int64(n), _p(seqs), C.c_int32(stride), _p(lens), _p(parent),
                                           C.byref(regs_p), _p(reg_off))
        self.bsq.check(rc, "bsq_align_phase1")
        total = int(reg_off[n])
        if total:
            buf = (C.c_char * (total * REG_DTYPE.itemsize)).from_address(regs_p.value)
            regs = np.frombuffer(buf, dtype=REG_DTYPE).copy()
        else:
            regs = np.zeros(0, dtype=REG_DTYPE)
        self.bsq.lib.bsq_free(regs_p)
        return regs, reg_off

    def counters(self) -> np.ndarray:
        c = np.zeros(16, dtype=np.int64)
        self.bsq.check(self.bsq.lib.bsq_aligner_counters(self.h, _p(c), C.c_int(16)), "bsq_aligner_counters")
        return c

    # staged execution with the deferred fetch (bsq_aligner_stage / _run / _result_slot / _fetch_slot / _release_slot)
    def stage_run(self, seqs: np.ndarray, lens: np.ndarray, parent: np.ndarray):
        """Stage and run one batch; returns (slot, n_tasks, n_regs) with the result slot claimed."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        parent = np.ascontiguousarray(parent, dtype=np.uint8)
        L = self.bsq.lib
        self.bsq.check(L.bsq_aligner_stage(self.h, C.c_int64(seqs.shape[0]), _p(seqs), C.c_int32(seqs.shape[1]), _p(lens), _p(parent)), "bsq_aligner_stage")
        n_regs = C.c_int64()
        self.bsq.check(L.bsq_aligner_run(self.h, C.byref(n_regs)), "bsq_aligner_run")
        slot, nt, nr = C.c_int(), C.c_int64(), C.c_int64()
        self.bsq.check(L.bsq_aligner_result_slot(self.h, C.byref(slot), C.byref(nt), C.byref(nr)), "bsq_aligner_result_slot")
        return slot.value, nt.value, nr.value

    def fetch_slot(self, slot: int, n_tasks: int, n_regs: int):
        regs = np.zeros(max(n_regs, 1), dtype=REG_DTYPE)
        reg_off = np.zeros(n_tasks + 1, dtype=np.int64)
        self.bsq.check(self.bsq.lib.bsq_aligner_fetch_slot(self.h, C.c_int(slot), _p(regs), _p(reg_off)), "bsq_aligner_fetch_slot")
        return regs[:n_regs], reg_off

    def release_slot(self, slot: int):
        self.bsq.check(self.bsq.lib.bsq_aligner_release_slot(self.h, C.c_int(slot)), "bsq_aligner_release_slot")


CIGAR_JOB_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("row", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("w", "<i4"), ("truesc", "<i4"),
                            ("clip5", "<i4"), ("clip3", "<i4"), ("parent", "u1"), ("pad_", "u1", (3,))])
CIGAR_RES_DTYPE = np.dtype([("n_cigar", "<i4"), ("NM", "<i4"), ("ZC", "<i4"), ("ZR", "<i4"), ("score", "<i4"), ("lead_del", "<i4"),
                            ("bss_u", "<i4"), ("off", "<u4")])
MATESW_JOB_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("row", "<i4"), ("xtra", "<i4"), ("use_ga", "u1"), ("pad_", "u1", (7,))])
MATESW_RES_DTYPE = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"),
                             ("qb", "<i4"), ("pad_", "<i4")])
assert CIGAR_JOB_DTYPE.itemsize == 48 and CIGAR_RES_DTYPE.itemsize == 32 and MATESW_JOB_DTYPE.itemsize == 32 and MATESW_RES_DTYPE.itemsize == 32


class Dp:
    """bsq_dp: the batched phase-2 dynamic programming (final CIGAR/MD, mate-rescue local alignment)."""

    def __init__(self, idx: DevIndex, opt: Opt):
        self.idx, self.bsq, self.opt = idx, idx.bsq, opt
        h = C.c_void_p()
        self.bsq.check(self.bsq.lib.bsq_dp_create(idx.h, C.byref(opt), C.byref(h)), "bsq_dp_create")
        self.h = h
        self._keep = None

    def close(self):
        if self.h:
            self.bsq.lib.bsq_dp_destroy(self.h)
            self.h = None

    def set_reads(self, seqs: np.ndarray, lens: np.ndarray):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        self._keep = (seqs, lens)
        self.bsq.check(self.bsq.lib.bsq_dp_set_reads(self.h, C.c_int64(seqs.shape[0]), _p(seqs), C.c_int32(seqs.shape[1]), _p(lens)),
                       "bsq_dp_set_reads")
        self.bsq.check(self.bsq.lib.bsq_dp_sync(self.h), "bsq_dp_sync")

    def cigar(self, jobs: np.ndarray):
        """Returns (results, blob as uint32 array)."""
        jobs = np.ascontiguousarray(jobs, dtype=CIGAR_JOB_DTYPE)
        res = np.zeros(len(jobs), dtype=CIGAR_RES_DTYPE)
        self.bsq.check(self.bsq.lib.bsq_dp_cigar_submit(self.h, C.c_int64(len(jobs)), _p(jobs), _p(res)), "bsq_dp_cigar_submit")
        blob_p, words = C.c_void_p(), C.c_int64()
        self.bsq.check(self.bsq.lib.bsq_dp_cigar_wait(self.h, C.byref(blob_p), C.byref(words)), "bsq_dp_cigar_wait")
        if words.value:
            blob = np.frombuffer((C.c_char * (words.value * 4)).from_address(blob_p.value), dtype=np.uint32).copy()
        else:
            blob = np.zeros(0, dtype=np.uint32)
        return res, blob

    def matesw(self, jobs: np.ndarray) -> np.ndarray:
        jobs = np.ascontiguousarray(jobs, dtype=MATESW_JOB_DTYPE)
        res = np.zeros(len(jobs), dtype=MATESW_RES_DTYPE)
        self.bsq.check(self.bsq.lib.bsq_dp_matesw_submit(self.h, C.c_int64(len(jobs)), _p(jobs), _p(res)), "bsq_dp_matesw_submit")
        self.bsq.check(self.bsq.lib.bsq_dp_matesw_wait(self.h), "bsq_dp_matesw_wait")
        return res

    def counters(self) -> np.ndarray:
        c = np.zeros(8, dtype=np.int64)
        self.bsq.check(self.bsq.lib.bsq_dp_counters(self.h, _p(c), C.c_int(8)), "bsq_dp_counters")
        return c


_default = None


def load() -> Bsq:
    """The product library.  Raises BsqError when libbsq.so has not been built."""
    global _default
    if _default is None:
        _default = Bsq(LIB_PATH)
    return _default
