"""Edge cases of the phase-1 ABI (empty and ragged inputs, maximum sizes): reads of the maximum
supported length (BSQ_MAX_READ_LEN = 256) with indels against the unmodified reference, an over-long read rejected
loudly, empty batches, and rows whose stride leaves them unaligned."""
import numpy as np
import pytest

import refprobe
import synth
from biscuit_b200 import capi, indexio

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def long_case(ds_hard):
    hi = indexio.load_index(ds_hard["fa"])
    rp = refprobe.RefProbe(ds_hard["fa"])
    p = synth.simulate_pairs(ds_hard["ref"], 60, seed=17, read_len=256, ins_mean=520.0, ins_sd=40.0, sub_rate=0.01, indel_rate=0.002)
    reads = [np.asarray(r, dtype=np.uint8)[:256] for r in list(p["r1"]) + list(p["r2"])]
    yield hi, rp, reads
    rp.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_max_length_reads(long_case, backend, request):
    hi, rp, reads = long_case
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    al = capi.Aligner(dx, bsq.default_opt())
    n = len(reads)
    stride = 259  # odd stride: most rows are not 8-byte aligned
    mat = np.zeros((n, stride), np.uint8)
    lens = np.array([len(r) for r in reads], np.int32)
    for i, r in enumerate(reads):
        mat[i, :len(r)] = r
    assert lens.max() == 256
    tasks = np.concatenate([mat, mat])
    tl = np.concatenate([lens, lens])
    par = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
    regs, off = al.phase1(tasks, tl, par)
    mine = refprobe.regs_from_bsq(regs)
    n_regs = 0
    for t in range(2 * n):
        exp = refprobe.regs_from_ref(rp.align1(int(par[t]), reads[t % n]))
        got = mine[off[t]:off[t + 1]]
        assert got.shape == exp.shape and (got == exp).all(), (t, got, exp)
        n_regs += len(exp)
    assert n_regs >= n
    al.close()
    dx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_overlong_read_rejected_and_empty_batch(long_case, backend, request):
    hi, rp, reads = long_case
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    al = capi.Aligner(dx, bsq.default_opt())
    mat = np.zeros((2, 272), np.uint8)
    mat[0, :257] = np.resize(reads[0], 257)
    mat[1, :100] = reads[1][:100]
    with pytest.raises(capi.BsqError):
        al.phase1(mat, np.array([257, 100], np.int32), np.array([0, 1], np.uint8))
    # the aligner is still usable afterwards, and an empty batch is a valid call
    regs, off = al.phase1(np.zeros((0, 160), np.uint8), np.zeros(0, np.int32), np.zeros(0, np.uint8))
    assert len(regs) == 0 and off.tolist() == [0]
    regs, off = al.phase1(mat[1:2], np.array([100], np.int32), np.array([1], np.uint8))
    exp = refprobe.regs_from_ref(rp.align1(1, reads[1][:100]))
    assert (refprobe.regs_from_bsq(regs) == exp).all()
    al.close()
    dx.close()


@pytest.fixture(scope="module")
def repeat_case(tmp_path_factory):
    """A reference with a 900-copy tandem repeat: SMEM intervals with more than max_occ (500) occurrences, chains
    with equal positions, tasks with thousands of seeds -- everything the shared-memory chaining declines and hands
    to the exact fallback."""
    import os
    import subprocess
    if not refprobe.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(99)
    unit = rng.integers(0, 4, size=700, dtype=np.uint8)
    rep = np.tile(unit, 900)
    mut = rng.random(len(rep)) < 0.002  # a few differences between copies
    rep[mut] = (rep[mut] + 1 + rng.integers(0, 3, size=int(mut.sum()))) % 4
    uniq = rng.integers(0, 4, size=120_000, dtype=np.uint8)
    ref = [("chrU", uniq), ("chrR", rep.astype(np.uint8))]
    d = tmp_path_factory.mktemp("repeat")
    fa = os.path.join(str(d), "ref.fa")
    synth.write_fasta(fa, ref)
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = synth.simulate_pairs(ref, 40, seed=5, sub_rate=0.01)
    reads = [np.asarray(r, dtype=np.uint8) for r in list(p["r1"]) + list(p["r2"])]
    hi = indexio.load_index(fa)
    rp = refprobe.RefProbe(fa)
    yield hi, rp, reads
    rp.close()


@pytest.mark.parametrize("fb_pool", [None, "2000"])  # "2000": the fallback workspace is too small at first -> grow and retry
@pytest.mark.parametrize("backend", BACKENDS)
def test_repeats_take_the_exact_fallback(repeat_case, backend, fb_pool, request, monkeypatch):
    hi, rp, reads = repeat_case
    if fb_pool:
        monkeypatch.setenv("BSQ_FB_POOL", fb_pool)
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    al = capi.Aligner(dx, bsq.default_opt())
    n = len(reads)
    mat = np.stack(reads)
    lens = np.full(n, mat.shape[1], np.int32)
    tasks = np.concatenate([mat, mat])
    tl = np.concatenate([lens, lens])
    par = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
    regs, off = al.phase1(tasks, tl, par)
    mine = refprobe.regs_from_bsq(regs)
    for t in range(2 * n):
        exp = refprobe.regs_from_ref(rp.align1(int(par[t]), reads[t % n]))
        got = mine[off[t]:off[t + 1]]
        assert got.shape == exp.shape and (got == exp).all(), (t, got, exp)
    c = al.counters()
    if backend == "cuda":
        assert c[14] > 0  # some tasks really went through k_chain
    al.close()
    dx.close()


def test_typed_introsort_matches_generic(tmp_path):
    """biscuit_b200/host/bq_sort.h (typed instances used by phase 2) against bq_introsort (the exact klib sequence): same
    order of equal keys on 6000 random / patterned arrays (tests/hostemu/sort_check.c)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "sort_check")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-o", exe, os.path.join(root, "tests", "hostemu", "sort_check.c"),
                           os.path.join(root, "biscuit_b200", "host", "bq_core.c"), "-lm"])
    out = subprocess.run([exe], stdout=subprocess.PIPE, check=True).stdout
    assert out.startswith(b"ok ")
