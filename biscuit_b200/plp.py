"""ctypes mirror of the pileup part of include/bsq.h (and, for tests, of oracle/bsq_oracle.h: same structs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

REC_DTYPE = np.dtype([("pos", "<i4"), ("dp", "<i4"), ("meth", "<i4", (3,)), ("base", "<i4", (7,)), ("base_redist", "<i4", (7,)),
                      ("rb_code", "u1"), ("cm1", "i1"), ("ctx", "u1"), ("methcallable", "u1"), ("n5", "S5"), ("any_callable", "u1"),
                      ("pad_", "u1", (2,))])
assert REC_DTYPE.itemsize == 88

_FIELDS = [("pos", "pos"), ("mpos", "mpos"), ("mate_rlen", "mate_rlen"), ("l_qseq", "l_qseq"), ("nm", "nm"), ("as_", "as"), ("flag", "flag"),
           ("mapq", "mapq"), ("bss_tag", "bss_tag"), ("sid", "sid"), ("n_cigar", "n_cigar"), ("cigar_off", "cigar_off"), ("cigar", "cigar"),
           ("seq_off", "seq_off"), ("seq", "seq"), ("qual_off", "qual_off"), ("qual", "qual")]


class Reads(C.Structure):
    _fields_ = [("n_reads", C.c_int64)] + [(c, C.c_void_p) for _, c in _FIELDS]


class Conf(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("min_base_qual", "min_read_len", "min_dist_end_5p", "min_dist_end_3p", "min_mapq", "min_score",
                                         "max_nm", "max_retention", "filter_ppair", "filter_secondary", "filter_duplicate", "filter_qcfail",
                                         "filter_doublecnt", "ambi_redist", "verbose", "is_nome")]


def make_reads_struct(rd: dict):
    """rd: dict of numpy arrays (tools/synth_plp.make_reads).  Returns (Reads, keepalive)."""
    r = Reads()
    r.n_reads = int(rd["n_reads"])
    keep = []
    for py, cname in _FIELDS:
        a = np.ascontiguousarray(rd[py])
        keep.append(a)
        setattr(r, cname, a.ctypes.data)
    return r, keep


class Pileup:
    """bsq_plp_* of the product library."""

    def __init__(self, bsq, n_bams: int = 1, device: int = 0):
        self.bsq, self.n_bams = bsq, n_bams
        h = C.c_void_p()
        bsq.check(bsq.lib.bsq_plp_create(C.c_int(device), C.c_int(n_bams), C.byref(h)), "bsq_plp_create")
        self.h = h
        bsq.lib.bsq_plp_destroy.argtypes = [C.c_void_p]

    def default_conf(self) -> Conf:
        c = Conf()
        self.bsq.lib.bsq_plp_conf_default(C.byref(c))
        return c

    def close(self):
        if self.h:
            self.bsq.lib.bsq_plp_destroy(self.h)
            self.h = None

    def set_contig(self, ref_nt4: np.ndarray):
        ref = np.ascontiguousarray(ref_nt4, dtype=np.uint8)
        self.bsq.check(self.bsq.lib.bsq_plp_set_contig(self.h, ref.ctypes.data_as(C.c_void_p), C.c_int32(len(ref))), "bsq_plp_set_contig")

    def stage(self, rd: dict):
        r, keep = make_reads_struct(rd)
        self.bsq.check(self.bsq.lib.bsq_plp_stage(self.h, C.byref(r)), "bsq_plp_stage")

    def run(self, conf: Conf, beg: int, end: int) -> int:
        n = C.c_int64()
        self.bsq.check(self.bsq.lib.bsq_plp_run(self.h, C.byref(conf), C.c_int32(beg), C.c_int32(end), C.byref(n)), "bsq_plp_run")
        return n.value

    def fetch(self, n_loci: int) -> np.ndarray:
        out = np.zeros(n_loci * self.n_bams, dtype=REC_DTYPE)
        if n_loci:
            self.bsq.check(self.bsq.lib.bsq_plp_fetch(self.h, out.ctypes.data_as(C.c_void_p)), "bsq_plp_fetch")
        return out

    def counters(self) -> np.ndarray:
        c = np.zeros(8, np.int64)
        self.bsq.check(self.bsq.lib.bsq_plp_counters(self.h, c.ctypes.data_as(C.c_void_p), C.c_int(8)), "bsq_plp_counters")
        return c

    def region(self, conf: Conf, rd: dict, beg: int, end: int) -> np.ndarray:
        self.stage(rd)
        return self.fetch(self.run(conf, beg, end))
