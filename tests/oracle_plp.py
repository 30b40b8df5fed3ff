"""ctypes access to oracle/libbsq_oracle.so (the CPU restatement) -- test infrastructure only."""
import ctypes as C
import os

import numpy as np

from biscuit_b200 import plp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libbsq_oracle.so")


def available():
    return os.path.exists(SO)


def conf_default() -> plp.Conf:
    lib = C.CDLL(SO)
    c = plp.Conf()
    lib.bsqo_plp_conf_default(C.byref(c))
    return c


def region(conf: plp.Conf, ref_nt4: np.ndarray, rd: dict, beg: int, end: int, n_bams: int = 1) -> np.ndarray:
    lib = C.CDLL(SO)
    lib.bsqo_plp_region.restype = C.c_int64
    r, keep = plp.make_reads_struct(rd)
    ref = np.ascontiguousarray(ref_nt4, dtype=np.uint8)
    cap = max(1, min(end, len(ref)) - max(beg, 1) + 1)
    out = np.zeros(cap * n_bams, dtype=plp.REC_DTYPE)
    n = lib.bsqo_plp_region(C.byref(conf), ref.ctypes.data_as(C.c_void_p), C.c_int32(len(ref)), C.c_int32(beg), C.c_int32(end), C.byref(r),
                            C.c_int(n_bams), out.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    assert n >= 0, n
    return out[: n * n_bams].copy()
