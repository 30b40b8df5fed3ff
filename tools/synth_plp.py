"""Synthetic coordinate-sorted bisulfite alignments for the pileup path, as the structure-of-arrays the
pileup ABI takes (= BAM record fields).  Reads come from tools/synth.simulate_pairs truth positions, so no
aligner is needed; `noise` sprinkles in the things the reference filters on (low MAPQ, duplicate /
secondary / QC-fail / improper flags, missing tags, indel / soft-clip CIGARs, missing strand tags)."""
import numpy as np

import synth

NT4_TO_NT16 = np.array([1, 2, 4, 8, 15], np.uint8)  # A C G T N in BAM's 4-bit code


def make_reads(contig_nt4: np.ndarray, n_pairs: int, seed: int = 5, noise: bool = True, n_bams: int = 1, read_len: int = 150):
    rng = np.random.default_rng(seed)
    p = synth.simulate_pairs([("c", contig_nt4)], n_pairs, seed=seed, read_len=read_len, qual="mixed" if noise else "const")
    _, pos, bsc, flen = p["truth"]
    rows = []
    for i in range(n_pairs):
        f0, fl = int(pos[i]), int(flen[i])
        r1, r2, q1, q2 = p["r1"][i], p["r2"][i], p["q1"][i], p["q2"][i]
        if not bsc[i]:  # BSW fragment: read 1 forward at the left end, read 2 reverse at the right end
            a = dict(pos=f0, seq=r1, qual=q1, flag=99, bss=0, mpos=f0 + fl - read_len)
            b = dict(pos=f0 + fl - read_len, seq=(3 - r2[::-1]) % 4 if False else _rc(r2), qual=q2[::-1], flag=147, bss=0, mpos=f0)
        else:  # BSC fragment: read 1 reverse at the right end, read 2 forward at the left end
            a = dict(pos=f0 + fl - read_len, seq=_rc(r1), qual=q1[::-1], flag=83, bss=1, mpos=f0)
            b = dict(pos=f0, seq=r2, qual=q2, flag=163, bss=1, mpos=f0 + fl - read_len)
        for r in (a, b):
            r["cigar"] = [(read_len, 0)]
            r["mapq"], r["nm"], r["as"], r["mrl"], r["sid"] = 60, 1, read_len - 4, read_len, int(rng.integers(0, n_bams))
            if noise:
                u = rng.random()
                if u < 0.05:
                    r["mapq"] = int(rng.integers(0, 40))
                elif u < 0.08:
                    r["flag"] |= 0x400
                elif u < 0.10:
                    r["flag"] |= 0x100
                elif u < 0.12:
                    r["flag"] |= 0x200
                elif u < 0.15:
                    r["flag"] &= ~0x2
                elif u < 0.18:
                    r["as"] = int(rng.integers(0, 60))
                elif u < 0.21:
                    r["nm"] = None
                    r["as"] = None
                elif u < 0.25:
                    r["mrl"] = -1
                elif u < 0.32:
                    r["bss"] = -1
                elif u < 0.42:  # indel / clip CIGARs with the same query length
                    k = int(rng.integers(0, 6))
                    L = read_len
                    if k == 0:
                        r["cigar"] = [(5, 4), (L - 5, 0)]
                    elif k == 1:
                        r["cigar"] = [(40, 0), (3, 1), (L - 43, 0)]
                    elif k == 2:
                        r["cigar"] = [(60, 0), (7, 2), (L - 70, 0), (10, 4)]
                    elif k == 3:  # leading hard clip: the reference advances qpos over it (pileup.c:822-824)
                        r["cigar"] = [(10, 5), (L, 0)]
                    elif k == 4:
                        r["cigar"] = [(L, 0), (25, 5)]
                    else:
                        r["cigar"] = [(20, 5), (50, 0), (2, 2), (L - 50, 0)]
            rows.append(r)
    rows.sort(key=lambda r: r["pos"])
    n = len(rows)
    out = dict(n_reads=n,
               pos=np.array([r["pos"] for r in rows], np.int32), mpos=np.array([r["mpos"] for r in rows], np.int32),
               mate_rlen=np.array([r["mrl"] for r in rows], np.int32), l_qseq=np.full(n, read_len, np.int32),
               nm=np.array([np.iinfo(np.int32).min if r["nm"] is None else r["nm"] for r in rows], np.int32),
               as_=np.array([np.iinfo(np.int32).min if r["as"] is None else r["as"] for r in rows], np.int32),
               flag=np.array([r["flag"] for r in rows], np.uint16), mapq=np.array([r["mapq"] for r in rows], np.uint8),
               bss_tag=np.array([r["bss"] for r in rows], np.int8), sid=np.array([r["sid"] for r in rows], np.uint8),
               n_cigar=np.array([len(r["cigar"]) for r in rows], np.int32))
    out["cigar_off"] = np.concatenate([[0], np.cumsum(out["n_cigar"])[:-1]]).astype(np.int64)
    out["cigar"] = np.array([(ln << 4) | op for r in rows for ln, op in r["cigar"]], np.uint32)
    nb = (read_len + 1) // 2
    seq = np.zeros((n, nb * 2), np.uint8)
    for i, r in enumerate(rows):
        seq[i, :read_len] = NT4_TO_NT16[np.minimum(r["seq"], 4)]
    out["seq"] = ((seq[:, 0::2] << 4) | seq[:, 1::2]).astype(np.uint8).reshape(-1)
    out["seq_off"] = (np.arange(n, dtype=np.int64) * nb)
    out["qual"] = np.concatenate([(r["qual"].astype(np.int16) - 33).astype(np.uint8) for r in rows])
    out["qual_off"] = (np.arange(n, dtype=np.int64) * read_len)
    return out


def _rc(a):
    a = np.asarray(a)
    return np.where(a[::-1] < 4, 3 - a[::-1], 4).astype(np.uint8)
