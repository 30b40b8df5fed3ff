"""Reader for the reference's on-disk index (`biscuit index` output), numpy only.

Layout (reference lib/aln/bwt.c:402-422, bntseq.c:68-216):
  <p>.{par,dau}.bwt : u64 primary, u64 L2[1..4], u32 bwt[]  (64-byte blocks: u64 occ[4] + 128 2-bit symbols)
  <p>.{par,dau}.sa  : u64 primary, u64 L2[1..4], u64 sa_intv, u64 seq_len, u64 sa[1..n_sa-1]
  <p>.bis.pac       : 2-bit packed forward reference + trailer byte(s)
  <p>.bis.ann/.amb  : text
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class FmHalf:
    primary: int
    L2: np.ndarray  # u64[5], L2[0] = 0
    bwt: np.ndarray  # u32 words, 64-byte aligned copy
    sa: np.ndarray  # u64[n_sa], sa[0] = 2**64-1
    sa_intv: int
    seq_len: int


@dataclasses.dataclass
class HostIndex:
    fm: list  # [daughter, parent]
    pac: np.ndarray
    l_pac: int
    names: list
    ann_offset: np.ndarray
    ann_len: np.ndarray
    ann_is_alt: np.ndarray


def _aligned_u32(n_words: int) -> np.ndarray:
    raw = np.zeros(n_words * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n_words * 4].view(np.uint32)


def _load_half(prefix: str, tag: str) -> FmHalf:
    with open(f"{prefix}.{tag}.bwt", "rb") as fh:
        hdr = np.frombuffer(fh.read(40), dtype=np.uint64)
        body = np.frombuffer(fh.read(), dtype=np.uint32)
    bwt = _aligned_u32(len(body))
    bwt[:] = body
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = hdr[1:5]
    with open(f"{prefix}.{tag}.sa", "rb") as fh:
        h2 = np.frombuffer(fh.read(56), dtype=np.uint64)
        rest = np.frombuffer(fh.read(), dtype=np.uint64)
    assert int(h2[0]) == int(hdr[0]), "SA-BWT inconsistency: primary"
    sa_intv, seq_len = int(h2[5]), int(h2[6])
    assert seq_len == int(L2[4]), "SA-BWT inconsistency: seq_len"
    n_sa = (seq_len + sa_intv) // sa_intv
    sa = np.empty(n_sa, dtype=np.uint64)
    sa[0] = np.uint64(2**64 - 1)
    sa[1:] = rest[: n_sa - 1]
    return FmHalf(int(hdr[0]), L2, bwt, sa, sa_intv, seq_len)


def load_index(prefix: str) -> HostIndex:
    fm = [_load_half(prefix, "dau"), _load_half(prefix, "par")]
    with open(f"{prefix}.bis.ann") as fh:
        toks = fh.read().split("\n")
    l_pac, n_seqs, _seed = (int(x) for x in toks[0].split())
    names, offs, lens = [], [], []
    for i in range(n_seqs):
        names.append(toks[1 + 2 * i].split()[1])
        o, ln, _n = toks[2 + 2 * i].split()
        offs.append(int(o))
        lens.append(int(ln))
    pac = np.fromfile(f"{prefix}.bis.pac", dtype=np.uint8)[: l_pac // 4 + 1].copy()
    alt = np.zeros(n_seqs, dtype=np.int32)
    try:
        with open(f"{prefix}.alt") as fh:
            altn = {ln.split("\t")[0].strip() for ln in fh if not ln.startswith("@")}
        for i, nm in enumerate(names):
            alt[i] = int(nm in altn)
    except FileNotFoundError:
        pass
    return HostIndex(fm, pac, l_pac, names, np.array(offs, dtype=np.int64), np.array(lens, dtype=np.int32), alt)
