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
