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

    def fetch_into(self, out: np.ndarray, n_loci: int) -> np.ndarray:
        """bsq_plp_fetch into a caller-owned buffer (e.g. page-locked memory): no allocation, no clearing."""
        assert out.dtype == REC_DTYPE and len(out) >= n_loci * self.n_bams
        if n_loci:
            self.bsq.check(self.bsq.lib.bsq_plp_fetch(self.h, out.ctypes.data_as(C.c_void_p)), "bsq_plp_fetch")
        return out[: n_loci * self.n_bams]

    def counters(self) -> np.ndarray:
        c = np.zeros(8, np.int64)
        self.bsq.check(self.bsq.lib.bsq_plp_counters(self.h, c.ctypes.data_as(C.c_void_p), C.c_int(8)), "bsq_plp_counters")
        return c

    def region(self, conf: Conf, rd: dict, beg: int, end: int) -> np.ndarray:
        self.stage(rd)
        return self.fetch(self.run(conf, beg, end))


# ---- multi-GPU: the only cross-rank state of the pileup path -------------------------------
def context_stats(recs: np.ndarray, n_bams: int, pos0: int = 1, step: int = 100000):
    """Per-sample methylation statistics of emitted records, as write_func accumulates them (src/pileup.c:178-185):
    cnt[sid, ctx] (int64) and betasum[sid, ctx] (float64) over the 6 cytosine contexts.  betasum is summed per
    window of `step` loci in locus order and the windows are then added in order, like the reference's records."""
    cnt = np.zeros((n_bams, 6), np.int64)
    beta = np.zeros((n_bams, 6), np.float64)
    if len(recs) == 0:
        return cnt, beta
    r = recs.reshape(-1, n_bams)
    win = (r["pos"][:, 0].astype(np.int64) - pos0) // step
    for s in range(n_bams):
        q = r[:, s]
        ok = (q["methcallable"] != 0) & (q["ctx"] < 6)
        ret = q["meth"][:, 0].astype(np.float64)
        tot = (q["meth"][:, 0] + q["meth"][:, 1]).astype(np.float64)
        b = np.divide(ret, tot, out=np.zeros_like(ret), where=ok)
        for ctx in range(6):
            m = ok & (q["ctx"] == ctx)
            cnt[s, ctx] = int(m.sum())
            acc = 0.0
            for w in np.unique(win[m]):
                part = 0.0
                for v in b[m & (win == w)]:  # locus order inside the window
                    part += float(v)
                acc += part
            beta[s, ctx] = acc
    return cnt, beta


def merge_stats(cnt: np.ndarray, beta: np.ndarray, device=None):
    """Final reduce across ranks (torch.distributed; NCCL over NVLink when `device` is a CUDA device, gloo on CPU).
    cnt / beta: [n_contigs_total, n_bams, 6] with zeros for the contigs this rank did not process.
    Integer counts: one all_reduce(SUM).  The double sums are all-gathered and added in rank order on every rank, so
    the `%1.3f` figures of <out>_meth_average.tsv do not depend on the reduction tree."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return cnt.copy(), beta.copy()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    c = torch.from_numpy(np.ascontiguousarray(cnt)).to(dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    b = torch.from_numpy(np.ascontiguousarray(beta)).to(dev)
    parts = [torch.empty_like(b) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, b)
    tot = torch.zeros_like(b)
    for p in parts:  # fixed order
        tot += p
    return c.cpu().numpy(), tot.cpu().numpy()
