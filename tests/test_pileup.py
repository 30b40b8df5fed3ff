"""Pileup parity: CUDA path (bsq_plp_*) vs the CPU restatement (oracle/bsq_oracle_pileup.c) -- bit-exact integer
counts and per-locus decisions (SURVEY.md §8a p1-p4, p6 and the integer part of p7).  The oracle itself is
checked on hand-computed cases (the reference has no golden vectors for this path: parity unpinned for the
floating-point VCF fields, which are not produced here)."""
import numpy as np
import pytest

import oracle_plp
import synth
import synth_plp
from biscuit_b200 import plp

I32MIN = np.iinfo(np.int32).min


def _mini_reads(entries, read_len):
    """entries: list of dict(pos, seq(str), flag, bss, cigar=[(len,op)], qual=int, mpos, mrl)"""
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
    n = len(entries)
    rd = dict(n_reads=n, pos=np.array([e["pos"] for e in entries], np.int32), mpos=np.array([e.get("mpos", 0) for e in entries], np.int32),
              mate_rlen=np.array([e.get("mrl", -1) for e in entries], np.int32), l_qseq=np.array([len(e["seq"]) for e in entries], np.int32),
              nm=np.full(n, I32MIN, np.int32), as_=np.full(n, I32MIN, np.int32), flag=np.array([e.get("flag", 0) for e in entries], np.uint16),
              mapq=np.full(n, 60, np.uint8), bss_tag=np.array([e.get("bss", 0) for e in entries], np.int8), sid=np.zeros(n, np.uint8),
              n_cigar=np.array([len(e["cigar"]) for e in entries], np.int32))
    rd["cigar_off"] = np.concatenate([[0], np.cumsum(rd["n_cigar"])[:-1]]).astype(np.int64)
    rd["cigar"] = np.array([(ln << 4) | op for e in entries for ln, op in e["cigar"]], np.uint32)
    seqs, soff, quals, qoff = [], [], [], []
    so = qo = 0
    for e in entries:
        c = [code[ch] for ch in e["seq"]] + [0]
        packed = [(c[i] << 4) | c[i + 1] for i in range(0, len(e["seq"]), 2)]
        seqs += packed
        soff.append(so)
        so += len(packed)
        quals += [e.get("qual", 40)] * len(e["seq"])
        qoff.append(qo)
        qo += len(e["seq"])
    rd["seq"], rd["seq_off"] = np.array(seqs, np.uint8), np.array(soff, np.int64)
    rd["qual"], rd["qual_off"] = np.array(quals, np.uint8), np.array(qoff, np.int64)
    return rd


def test_oracle_hand_case():
    """Reference (1-based):  1 A 2 A 3 C 4 G 5 T 6 A 7 C 8 C 9 G 10 T 11 A 12 A 13 A ...  Two BSW reads and one BSC read."""
    if not oracle_plp.available():
        pytest.skip("oracle not built")
    ref = np.array([0, 0, 1, 2, 3, 0, 1, 1, 2, 3] + [0] * 10, np.uint8)
    L = 12
    # BSW read at pos 0: AACGTATCGTAA  -> C3 retained, C7 converted (T), C8 retained
    # BSW read at pos 0: AATGTACTGTAA  -> C3 converted, C7 retained, C8 converted
    # BSC read at pos 0: AACATACCGTAA  -> G4 converted (A), G9 retained
    rd = _mini_reads([dict(pos=0, seq="AACGTATCGTAA", bss=0, cigar=[(L, 0)]), dict(pos=0, seq="AATGTACTGTAA", bss=0, cigar=[(L, 0)]),
                      dict(pos=0, seq="AACATACCGTAA", bss=1, cigar=[(L, 0)])], L)
    c = oracle_plp.conf_default()
    out = oracle_plp.region(c, ref, rd, 1, len(ref))
    by = {int(r["pos"]): r for r in out}
    # positions 1-3 and 10-12 are within 3 bases of a read end (qpos <= 3 or rlen < qpos + 3): filtered from the counts
    assert 3 not in by
    # C7: read 1 converted, read 2 retained, the BSC read is NA on a C; context T A [C] C G = HCHG
    assert by[7]["meth"].tolist() == [1, 1, 1] and by[7]["dp"] == 3 and by[7]["n5"] == b"TACCG" and by[7]["ctx"] == 1
    assert by[8]["meth"].tolist() == [1, 1, 1] and by[8]["ctx"] == 0 and by[8]["n5"] == b"ACCGT"  # CpG
    assert by[4]["meth"].tolist() == [0, 1, 2] and by[4]["rb_code"] == 2  # G4: one BSC conversion, BSW reads give NA
    assert by[9]["meth"].tolist() == [1, 0, 2] and by[9]["n5"] == b"TACGG" and by[9]["ctx"] == 0  # G9 seen from the C strand
    # T5: the BSW T calls are ambiguous (Y) but are redistributed to the reference T: no mutant, no methylation -> not emitted
    assert 5 not in by and 6 not in by


def test_oracle_hand_case_ctx_fix():
    """CHG vs CHH classification on the same hand case (position 7: C followed by C then G = CHG)."""
    if not oracle_plp.available():
        pytest.skip("oracle not built")
    ref = np.array([0, 0, 1, 2, 3, 0, 1, 1, 2, 3] + [0] * 10, np.uint8)
    rd = _mini_reads([dict(pos=0, seq="AACGTATCGTAA", bss=0, cigar=[(12, 0)])], 12)
    out = oracle_plp.region(oracle_plp.conf_default(), ref, rd, 1, len(ref))
    by = {int(r["pos"]): r for r in out}
    assert by[7]["ctx"] == 1  # HCHG


def test_oracle_mate_overlap_and_filters():
    if not oracle_plp.available():
        pytest.skip("oracle not built")
    ref = np.tile(np.array([1, 2, 0, 3], np.uint8), 50)
    s = "CGAT" * 10
    base = dict(seq=s, bss=0, cigar=[(40, 0)])
    rd = _mini_reads([dict(base, pos=0, flag=99, mpos=20, mrl=40), dict(base, pos=20, flag=147, mpos=0, mrl=40)], 40)
    c = oracle_plp.conf_default()
    out = oracle_plp.region(c, ref, rd, 1, len(ref))
    dp = {int(r["pos"]): int(r["dp"]) for r in out}
    assert max(dp.values()) == 1  # read-2 bases inside the overlap [21,40] are skipped
    c.filter_doublecnt = 0
    out = oracle_plp.region(c, ref, rd, 1, len(ref))
    assert max(int(r["dp"]) for r in out) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("n_bams,noise", [(1, False), (1, True), (2, True)])
def test_pileup_matches_oracle(cuda, n_bams, noise):
    ref = synth.make_reference(300_000, 1, seed=3, n_runs=3)[0][1]
    rd = synth_plp.make_reads(ref, 6000, seed=9, noise=noise, n_bams=n_bams)
    pl = plp.Pileup(cuda, n_bams)
    conf = pl.default_conf()
    pl.set_contig(ref)
    for beg, end in ((1, len(ref) + 5), (1000, 1001), (5000, 105000), (299990, 300001)):
        got = pl.region(conf, rd, beg, end)
        exp = oracle_plp.region(conf, ref, rd, beg, end, n_bams)
        assert len(got) == len(exp), (beg, end, len(got), len(exp))
        assert got.tobytes() == exp.tobytes(), (beg, end)
    # option switches
    conf.filter_doublecnt = 0
    conf.ambi_redist = 0
    conf.min_base_qual = 30
    conf.filter_ppair = 0
    got = pl.region(conf, rd, 1, len(ref))
    exp = oracle_plp.region(conf, ref, rd, 1, len(ref), n_bams)
    assert got.tobytes() == exp.tobytes()
    c = pl.counters()
    assert c[2] == len(exp) // n_bams and c[3] > 0
    pl.close()


@pytest.mark.gpu
def test_pileup_unsorted_and_empty(cuda):
    ref = synth.make_reference(50_000, 1, seed=4)[0][1]
    rd = synth_plp.make_reads(ref, 500, seed=2, noise=True)
    perm = np.random.default_rng(0).permutation(rd["n_reads"])
    rd2 = dict(rd)
    for k in ("pos", "mpos", "mate_rlen", "l_qseq", "nm", "as_", "flag", "mapq", "bss_tag", "sid", "n_cigar", "cigar_off", "seq_off", "qual_off"):
        rd2[k] = rd[k][perm]
    pl = plp.Pileup(cuda, 1)
    conf = pl.default_conf()
    pl.set_contig(ref)
    a = pl.region(conf, rd, 1, len(ref))
    b = pl.region(conf, rd2, 1, len(ref))
    assert a.tobytes() == b.tobytes()  # counts do not depend on read order
    empty = dict(rd)
    empty["n_reads"] = 0
    assert len(pl.region(conf, empty, 1, len(ref))) == 0
    pl.close()
