"""Size-independent checks of a device-resident FM-index pair against the text it was built from, usable at sizes
where the reference's own `biscuit index` (hours at 3.1 Gb) cannot be run: sampled ranks -- half of them above 2^32
when the doubled text is that long -- must (1) come in suffix order, (2) obey SA[LF(k)] = SA[k] - 1 with LF computed
from occ4 and the symbol totals (bwt_invPsi, lib/aln/bwt.c:54-60), (3) reproduce the symbol totals L2.

Used by tests/test_index_build.py (2L > 2^32) and by bench.py on the 3.1 Gb bench index before anything is timed."""
import numpy as np


def text_at(nt4: np.ndarray, which: int, pos: int, n: int) -> np.ndarray:
    """T[pos:pos+n] of the converted, doubled text (bis_bns_fasta2bntseq, lib/aln/bntseq.c:588-600): forward strand then
    reverse complement, C->T for the parent index (which=1), G->A for the daughter (which=0); clipped at 2L."""
    L = len(nt4)
    end = min(pos + n, 2 * L)
    parts = []
    if pos < L:
        parts.append(nt4[pos:min(end, L)])
    if end > L:
        a, b = max(pos, L), end  # T[p] = 3 - nt4[2L-1-p]
        seg = nt4[2 * L - b:2 * L - a]
        parts.append((3 - seg[::-1]).astype(np.uint8))
    t = np.concatenate(parts) if len(parts) > 1 else parts[0].copy()
    if which == 1:
        t[t == 1] = 3
    else:
        t[t == 2] = 0
    return t


def check_index(dx, nt4: np.ndarray, n_samples: int = 2000, seed: int = 1, n_order: int = 400, totals: bool = True) -> dict:
    L = len(nt4)
    n = 2 * L
    sz = dx.sizes()
    rng = np.random.default_rng(seed)
    cnt = None
    if totals:
        cnt = np.zeros(4, np.int64)
        for o in range(0, L, 1 << 28):
            cnt += np.bincount(nt4[o:o + (1 << 28)], minlength=4)[:4]
    out = {"ranks_above_2p32": 0, "order_pairs": 0, "lf_checked": 0, "ok": True}
    for which in (0, 1):
        L2 = sz["L2"][which].astype(np.int64)
        primary = int(sz["primary"][which])
        if cnt is not None:  # symbol totals of the doubled, converted text
            a, c, g, t = cnt
            tot = np.array([a + t, c + g, c + g, a + t], np.int64)
            if which == 1:
                tot = np.array([a + t, 0, c + g, a + t + c + g], np.int64)
            else:
                tot = np.array([a + t + c + g, c + g, 0, a + t], np.int64)
            assert (np.cumsum(tot) == L2[1:]).all(), "L2"
        lo = rng.integers(1, min(n, 1 << 32) - 2, size=n_samples // 2).astype(np.uint64)
        hi = rng.integers(1 << 32, n - 2, size=n_samples - n_samples // 2).astype(np.uint64) if n > (1 << 32) + 1000 else lo[:0]
        k = np.concatenate([lo, hi])
        out["ranks_above_2p32"] += int(len(hi))
        pa = dx.sa_lookup(which, k).astype(np.int64)
        pb = dx.sa_lookup(which, k + np.uint64(1)).astype(np.int64)
        assert ((pa >= 0) & (pa < n)).all(), "SA range"
        pick = np.concatenate([np.arange(min(n_order // 2, len(lo))), len(lo) + np.arange(min(n_order // 2, len(hi)))])
        for i in pick:  # consecutive ranks hold suffixes in increasing order
            a_, b_ = int(pa[i]), int(pb[i])
            sa_, sb_ = text_at(nt4, which, a_, 200).tobytes(), text_at(nt4, which, b_, 200).tobytes()
            assert sa_ < sb_ or (sa_ == sb_[:len(sa_)] and len(sa_) < len(sb_)), ("suffix order", which, int(k[i]), a_, b_)
        out["order_pairs"] += len(pick)
        # LF: the symbol in front of suffix SA[k] and its rank among equal symbols give the rank of suffix SA[k]-1
        ok = (pa > 0) & (k.astype(np.int64) != primary)
        kk, pp = k[ok], pa[ok]
        c = np.array([int(text_at(nt4, which, int(p) - 1, 1)[0]) for p in pp])
        occ = dx.occ4(which, kk).astype(np.int64)
        lf = L2[c] + occ[np.arange(len(kk)), c]
        prev = dx.sa_lookup(which, lf.astype(np.uint64)).astype(np.int64)
        assert (prev == pp - 1).all(), ("LF inversion", which)
        out["lf_checked"] += int(len(kk))
        assert int(dx.sa_lookup(which, np.array([primary], np.uint64))[0]) == 0, "primary"
    return out
