"""GPU index construction (bsq_index_build) against the reference's `biscuit index` output, byte for byte,
plus size-independent properties at a larger size (LF-walk inverts the BWT; SA samples are sorted)."""
import numpy as np
import pytest

import pack
import refprobe
import synth
from biscuit_b200 import indexio


@pytest.mark.gpu
@pytest.mark.parametrize("full_sa", ["0", "1"])
@pytest.mark.parametrize("ds_name", ["ds_1m", "ds_hard"])
def test_index_matches_reference(cuda, ds_name, full_sa, request, monkeypatch):
    ds = request.getfixturevalue(ds_name)
    hi = indexio.load_index(ds["fa"])
    monkeypatch.setenv("BSQ_FULL_SA", full_sa)
    dx = cuda.build_index(hi.pac, hi.l_pac, hi.names, hi.ann_offset, hi.ann_len)
    sz = dx.sizes()
    if refprobe.available():  # every rank's SA value (the builder's full SA, or the LF walk) against bwt_sa of the reference
        rp = refprobe.RefProbe(ds["fa"])
        rng = np.random.default_rng(5)
        for which in (0, 1):
            k = rng.integers(1, hi.fm[which].seq_len + 1, size=20000).astype(np.uint64)
            k[:3] = [0, hi.fm[which].primary, hi.fm[which].seq_len]
            assert (dx.sa_lookup(which, k) == rp.sa(which, k)).all()
        rp.close()
    for which in (0, 1):
        f = hi.fm[which]
        assert int(sz["primary"][which]) == f.primary
        assert (sz["L2"][which] == f.L2).all()
        bwt, sa = dx.download(which)
        assert len(bwt) == len(f.bwt) and (bwt == f.bwt).all()
        assert len(sa) == len(f.sa) and (sa == f.sa).all()
    dx.close()


@pytest.mark.gpu
def test_index_many_chunks(cuda, ds_1m, monkeypatch):
    """Force the bucketed multi-pass path (as used for GRCh38-sized references) on a small reference."""
    monkeypatch.setenv("BSQ_INDEX_CHUNK", "150000")
    hi = indexio.load_index(ds_1m["fa"])
    dx = cuda.build_index(hi.pac, hi.l_pac, hi.names, hi.ann_offset, hi.ann_len)
    assert dx.sizes()["stats"][0] > 10
    for which in (0, 1):
        bwt, sa = dx.download(which)
        assert (bwt == hi.fm[which].bwt).all() and (sa == hi.fm[which].sa).all()
    dx.close()


@pytest.mark.gpu
def test_index_low_complexity(cuda, tmp_path):
    """Tandem repeats and homopolymer runs force many tie-refinement passes."""
    import subprocess
    rng = np.random.default_rng(3)
    unit = rng.integers(0, 4, size=37).astype(np.uint8)
    seq = np.concatenate([rng.integers(0, 4, size=5000).astype(np.uint8), np.tile(unit, 200), np.zeros(3000, np.uint8),
                          rng.integers(0, 4, size=5000).astype(np.uint8), np.tile(unit, 100), np.full(2000, 3, np.uint8)])
    # the doubled text then ends in a run of A (revcomp of the leading T run): exhausted-suffix ties
    seq = np.concatenate([np.full(40, 3, np.uint8), seq])
    fa = str(tmp_path / "rep.fa")
    synth.write_fasta(fa, [("rep", seq)])
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    hi = indexio.load_index(fa)
    dx = cuda.build_index(hi.pac, hi.l_pac, hi.names, hi.ann_offset, hi.ann_len)
    for which in (0, 1):
        bwt, sa = dx.download(which)
        assert (bwt == hi.fm[which].bwt).all() and (sa == hi.fm[which].sa).all()
    assert dx.sizes()["stats"][1] > 10  # refinement really iterated
    dx.close()


@pytest.mark.gpu
def test_index_properties_large(cuda):
    """64 Mb reference (too slow for the CPU reference in a test): bsq_sa_lookup over the built index must
    return, for consecutive ranks, suffixes in increasing lexicographic order, and occ4 must be consistent with L2."""
    L = 64_000_000
    ref = synth.make_reference(L, 4, seed=11)
    nt4 = np.concatenate([s for _, s in ref])
    pac = pack.pack_pac(nt4)
    lens = np.array([len(s) for _, s in ref], np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    dx = cuda.build_index(pac, L, [n for n, _ in ref], offs, lens)
    sz = dx.sizes()
    n = 2 * L
    rng = np.random.default_rng(1)
    for which in (0, 1):
        # converted, doubled text
        fwd = nt4.copy()
        rc = (3 - nt4[::-1]).astype(np.uint8)
        T = np.concatenate([fwd, rc])
        if which == 1:
            T[T == 1] = 3
        else:
            T[T == 2] = 0
        assert (np.bincount(T, minlength=4).cumsum() == sz["L2"][which][1:]).all()
        r0 = rng.integers(1, n - 1, size=3000).astype(np.uint64)
        pa = dx.sa_lookup(which, r0).astype(np.int64)
        pb = dx.sa_lookup(which, r0 + np.uint64(1)).astype(np.int64)
        for a, b in zip(pa[:600], pb[:600]):
            sa_, sb_ = T[a:a + 200].tobytes(), T[b:b + 200].tobytes()
            assert sa_ < sb_ or (sa_ == sb_[:len(sa_)] and len(sa_) < len(sb_)), (a, b)
        # occ4 at the end of the text equals the symbol totals
        cnt = dx.occ4(which, np.array([n], np.uint64))[0]
        assert (cnt == np.bincount(T, minlength=4)).all()
        # the rank of text position 0 is the primary
        assert dx.sa_lookup(which, np.array([sz["primary"][which]], np.uint64))[0] == 0
    dx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("full_sa", ["0", "1"])
def test_index_beyond_2p32(cuda, full_sa, monkeypatch):
    """Doubled text longer than 2^32 symbols (L = 2.2 Gb, the regime of the GRCh38-sized bench index: ranks, `primary`
    corrections and SA entries that no longer fit 32 bits).  Sampled ranks, half of them above 2^32: suffix order of
    consecutive ranks, SA[LF(k)] = SA[k] - 1 with LF from occ4 + L2 (bwt_invPsi, lib/aln/bwt.c:54-60), symbol totals.
    full_sa=0 answers bsq_sa_lookup by the LF walk over the sampled SA (bwt_sa, bwt.c:87-97), full_sa=1 from the
    builder's full suffix array."""
    import indexcheck
    L = 2_200_000_000
    rng = np.random.default_rng(13)
    nt4 = rng.integers(0, 4, size=L, dtype=np.uint8)
    n_contigs = 12
    base = L // n_contigs
    lens = np.full(n_contigs, base, np.int32)
    lens[-1] = L - base * (n_contigs - 1)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    pac = pack.pack_pac(nt4)
    monkeypatch.setenv("BSQ_FULL_SA", full_sa)
    dx = cuda.build_index(pac, L, [f"c{i}" for i in range(n_contigs)], offs, lens)
    r = indexcheck.check_index(dx, nt4, n_samples=3000, seed=3, totals=(full_sa == "0"))
    assert r["ranks_above_2p32"] >= 2000 and r["lf_checked"] >= 5000 and r["order_pairs"] >= 700
    dx.close()
