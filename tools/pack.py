"""2-bit packing helpers shared by tests and bench.py (layout of lib/aln/bntseq.c:233-234)."""
import numpy as np


def pack_pac(nt4: np.ndarray) -> np.ndarray:
    """nt4 codes (0..3) -> .bis.pac body: base l in pac[l>>2] >> ((~l&3)<<1)."""
    n = len(nt4)
    pad = (-n) % 4
    a = np.concatenate([nt4.astype(np.uint8), np.zeros(pad, np.uint8)]).reshape(-1, 4)
    out = (a[:, 0] << 6) | (a[:, 1] << 4) | (a[:, 2] << 2) | a[:, 3]
    return np.concatenate([out.astype(np.uint8), np.zeros(1, np.uint8)])
