"""Deterministic synthetic inputs: iid-uniform reference contigs and simulated
2x150 bp bisulfite read pairs.  Used by tests, tools/make_golden.py and bench.py; pure numpy.

nt4 code: A0 C1 G2 T3 N4 (reference lib/aln/bntseq.c:49-66).
"""
from __future__ import annotations

import numpy as np

NT = np.frombuffer(b"ACGTN", dtype=np.uint8)


def make_reference(total_len: int, n_contigs: int = 1, seed: int = 7, n_runs: int = 0):
    """Return [(name, nt4 uint8 array)].  Contig lengths are equal except the last.
    n_runs > 0 plants that many short runs of N (code 4) to exercise the lrand48 path
    (reference lib/aln/bntseq.c:495,558-559)."""
    rng = np.random.default_rng(seed)
    base = total_len // n_contigs
    out = []
    for i in range(n_contigs):
        ln = base if i < n_contigs - 1 else total_len - base * (n_contigs - 1)
        seq = rng.integers(0, 4, size=ln, dtype=np.uint8)
        for _ in range(n_runs):
            p = int(rng.integers(0, max(1, ln - 50)))
            seq[p:p + int(rng.integers(1, 40))] = 4
        out.append((f"chr{i + 1}", seq))
    return out


def write_fasta(path: str, contigs, width: int = 100) -> None:
    with open(path, "wb") as fh:
        for name, seq in contigs:
            fh.write(b">" + name.encode() + b"\n")
            asc = NT[seq]
            n_full = len(asc) // width
            if n_full:
                body = np.empty((n_full, width + 1), dtype=np.uint8)
                body[:, :width] = asc[: n_full * width].reshape(n_full, width)
                body[:, width] = 10
                fh.write(body.tobytes())
            if len(asc) % width:
                fh.write(asc[n_full * width:].tobytes() + b"\n")


def simulate_pairs(contigs, n_pairs: int, seed: int = 1, read_len: int = 150, ins_mean: float = 300.0,
                   ins_sd: float = 30.0, sub_rate: float = 0.005, cpg_ret: float = 0.8, cph_ret: float = 0.01,
                   indel_rate: float = 0.0, qual: str = "const", n_rate: float = 0.0):
    """Simulate directional+complementary bisulfite pairs.

    Returns dict(r1, r2: list of nt4 arrays (ragged if indels) or (n,read_len) arrays,
    q1, q2: uint8 phred+33 arrays, truth: (contig, pos, is_bsc, frag_len)).
    """
    rng = np.random.default_rng(seed)
    lens = np.array([len(s) for _, s in contigs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    G = np.concatenate([s for _, s in contigs])
    G = np.where(G > 3, 0, G).astype(np.uint8)  # simulated molecules carry no N
    flen = np.clip(np.rint(rng.normal(ins_mean, ins_sd, n_pairs)), read_len + 10, None).astype(np.int64)
    M = int(flen.max())
    cid = rng.choice(len(contigs), size=n_pairs, p=lens / lens.sum())
    span = np.maximum(lens[cid] - flen - 2, 1)
    pos = (rng.random(n_pairs) * span).astype(np.int64) + 1
    start = offs[cid] + pos
    bsc = rng.random(n_pairs) < 0.5
    ar = np.arange(M + 1, dtype=np.int64)[None, :]
    idx = np.where(bsc[:, None], start[:, None] + flen[:, None] - 1 - ar, start[:, None] + ar)
    np.clip(idx, 0, len(G) - 1, out=idx)
    F = G[idx]
    F = np.where(bsc[:, None], 3 - F, F).astype(np.uint8)
    is_c = F[:, :M] == 1
    is_cpg = is_c & (F[:, 1:M + 1] == 2)
    keep_p = np.where(is_cpg, cpg_ret, cph_ret)
    conv = is_c & (rng.random((n_pairs, M)) >= keep_p)
    F = F[:, :M].copy()
    F[conv] = 3
    r1 = F[:, :read_len].copy()
    j = np.arange(read_len, dtype=np.int64)[None, :]
    r2 = (3 - np.take_along_axis(F, flen[:, None] - 1 - j, axis=1)).astype(np.uint8)

    def mutate(r):
        e = rng.random(r.shape) < sub_rate
        r[e] = (r[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) % 4
        if n_rate > 0:
            r[rng.random(r.shape) < n_rate] = 4
        return r

    r1 = mutate(r1)
    r2 = mutate(r2)
    if qual == "const":
        q1 = np.full(r1.shape, ord("I"), dtype=np.uint8)
        q2 = np.full(r2.shape, ord("I"), dtype=np.uint8)
    else:  # mixed: Phred ~ U[2,40]
        q1 = (33 + rng.integers(2, 41, size=r1.shape)).astype(np.uint8)
        q2 = (33 + rng.integers(2, 41, size=r2.shape)).astype(np.uint8)
    out = dict(r1=r1, r2=r2, q1=q1, q2=q2, truth=(cid, pos, bsc, flen))
    if indel_rate > 0:
        out["r1"], out["q1"] = _indel(rng, r1, q1, indel_rate, read_len)
        out["r2"], out["q2"] = _indel(rng, r2, q2, indel_rate, read_len)
    return out


def _indel(rng, reads, quals, rate, read_len):
    """Apply small insertions/deletions (length 1-3) and trim/pad back to read_len."""
    out_r = reads.copy()
    out_q = quals.copy()
    n_ev = rng.binomial(read_len, rate, size=len(reads))
    for i in np.nonzero(n_ev)[0]:
        r = list(reads[i])
        q = list(quals[i])
        for _ in range(int(n_ev[i])):
            p = int(rng.integers(5, len(r) - 5))
            ln = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                del r[p:p + ln]
                del q[p:p + ln]
            else:
                ins = rng.integers(0, 4, size=ln).tolist()
                r[p:p] = ins
                q[p:p] = [q[p]] * ln
        while len(r) < read_len:  # pad with random bases (acts like adaptor read-through)
            r.append(int(rng.integers(0, 4)))
            q.append(q[-1])
        out_r[i] = np.array(r[:read_len], dtype=np.uint8)
        out_q[i] = np.array(q[:read_len], dtype=np.uint8)
    return out_r, out_q


def write_fastq(path: str, reads, quals, first_id: int = 0, suffix: str = "") -> None:
    n, L = reads.shape
    asc = NT[reads]
    with open(path, "wb") as fh:
        chunk = []
        for i in range(n):
            chunk.append(b"@r%d%s\n" % (first_id + i, suffix.encode()))
            chunk.append(asc[i].tobytes())
            chunk.append(b"\n+\n")
            chunk.append(quals[i].tobytes())
            chunk.append(b"\n")
            if len(chunk) > 50000:
                fh.write(b"".join(chunk))
                chunk = []
        fh.write(b"".join(chunk))
