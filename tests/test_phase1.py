"""Parity of phase 1 of the aligner (seeding, SA lookup, chaining, chain filter, extension) against
the UNMODIFIED reference (oracle/_ref).  Bit-exact: all integer work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import refprobe
from biscuit_b200 import capi, indexio

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _reads(ds, extra_short=True):
    p = ds["pairs"]
    reads = [np.asarray(r, dtype=np.uint8) for r in p["r1"]] + [np.asarray(r, dtype=np.uint8) for r in p["r2"]]
    if extra_short:  # ragged inputs: short, sub-seed-length and single-base reads
        reads += [reads[0][:40], reads[1][:18], reads[2][:19], reads[3][:1], np.full(30, 4, np.uint8), reads[4][:75]]
    L = max(len(r) for r in reads)
    mat = np.zeros((len(reads), L), dtype=np.uint8)
    for i, r in enumerate(reads):
        mat[i, :len(r)] = r
    return reads, mat, np.array([len(r) for r in reads], dtype=np.int32)


@pytest.fixture(scope="module", params=["ds_1m", "ds_hard"])
def ctx(request):
    ds = request.getfixturevalue(request.param)
    hi = indexio.load_index(ds["fa"])
    rp = refprobe.RefProbe(ds["fa"])
    yield ds, hi, rp
    rp.close()


@pytest.mark.parametrize("full_sa", ["0", "1"])  # LF walk over the sampled SA vs. the full SA derived in HBM
@pytest.mark.parametrize("backend", BACKENDS)
def test_occ4_and_sa(ctx, backend, full_sa, request, monkeypatch):
    ds, hi, rp = ctx
    monkeypatch.setenv("BSQ_FULL_SA", full_sa)
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    rng = np.random.default_rng(0)
    for which in (0, 1):
        f = hi.fm[which]
        k = rng.integers(0, f.seq_len + 1, size=4000).astype(np.uint64)
        k[:6] = [2**64 - 1, f.primary, f.seq_len, 0, f.primary - 1, f.primary + 1]
        assert (dx.occ4(which, k) == rp.occ4(which, k.view(np.int64))).all()
        k[0] = 1
        assert (dx.sa_lookup(which, k) == rp.sa(which, k)).all()
    dx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_collect_intv(ctx, backend, request):
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    opt = bsq.default_opt()
    reads, mat, lens = _reads(ds)
    for parent in (0, 1):
        out, n_out = dx.collect_intv(opt, mat, lens, np.full(len(reads), parent))
        for i, r in enumerate(reads):
            exp = rp.collect_intv(parent, r)
            assert n_out[i] == len(exp), (i, parent)
            assert (out[i, :n_out[i]] == exp).all(), (i, parent)
    dx.close()


def test_chain_hostemu(ctx, hostemu):
    """mem_chain + mem_chain_flt (incl. B-tree and introsort tie order) through the host emulation hook."""
    ds, hi, rp = ctx
    dx = hostemu.upload(hi)
    opt = hostemu.default_opt()
    reads, _, _ = _reads(ds)
    hostemu.lib.hostemu_chain.restype = C.c_int64
    for parent in (0, 1):
        for i, r in enumerate(reads):
            exp, nch, fr = rp.chain(parent, r, stage=1)
            out = np.zeros(max(len(exp), 8) + 64, dtype=np.int64)
            nc, f = C.c_int(), C.c_float()
            seq = np.ascontiguousarray(r)
            o = hostemu.lib.hostemu_chain(dx.h, C.byref(opt), C.c_int(parent), C.c_int(len(r)), seq.ctypes.data_as(C.c_void_p),
                                          C.byref(nc), C.byref(f), out.ctypes.data_as(C.c_void_p), C.c_int64(len(out)))
            assert o == len(exp) and nc.value == nch, (i, parent)
            assert (out[:o] == exp).all(), (i, parent)
            if nch:
                assert f.value == fr
    dx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_phase1_regions(ctx, backend, request):
    """mem_align1_core: regions (before mem_merge_regions) for every read x conversion."""
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    opt = bsq.default_opt()
    al = capi.Aligner(dx, opt)
    reads, mat, lens = _reads(ds)
    n = len(reads)
    tasks = np.concatenate([mat, mat])
    tl = np.concatenate([lens, lens])
    par = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
    regs, off = al.phase1(tasks, tl, par)
    mine = refprobe.regs_from_bsq(regs)
    n_regs = 0
    for t in range(2 * n):
        exp = refprobe.regs_from_ref(rp.align1(int(par[t]), reads[t % n]))
        got = mine[off[t]:off[t + 1]]
        assert got.shape == exp.shape, (t, got, exp)
        assert (got == exp).all(), (t, got, exp)
        n_regs += len(exp)
    assert n_regs > n  # the data set really produced alignments
    al.close()
    dx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_extend_batch(ctx, backend, request):
    """ksw_extend2 on random jobs incl. indels, narrow bands, z-drop and zero-length targets."""
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    opt = bsq.default_opt()
    rng = np.random.default_rng(5)
    qs, ts, par, ws, h0s = [], [], [], [], []
    for j in range(300):
        ql = int(rng.integers(1, 151))
        q = rng.integers(0, 4, size=ql).astype(np.uint8)
        t = list(q)
        for _ in range(int(rng.integers(0, 4))):  # mutate the target: substitutions / indels
            p = int(rng.integers(0, len(t) + 1))
            kind = rng.integers(0, 3)
            if kind == 0 and p < len(t):
                t[p] = int(rng.integers(0, 4))
            elif kind == 1:
                t[p:p] = rng.integers(0, 4, size=int(rng.integers(1, 6))).tolist()
            elif p < len(t):
                del t[p:p + int(rng.integers(1, 6))]
        t = np.array(t + rng.integers(0, 4, size=int(rng.integers(0, 120))).tolist(), dtype=np.uint8)
        if j % 37 == 0:
            t = t[:0]
        if j % 11 == 0:
            q[rng.integers(0, ql)] = 4
        qs.append(q)
        ts.append(t)
        par.append(j & 1)
        ws.append(int(rng.choice([100, 200, 5, 1])))
        h0s.append(int(rng.integers(1, 151)))
    got = capi.extend_batch(bsq, opt, qs, ts, par, ws, h0s)
    for j in range(len(qs)):
        mat = np.array(opt.ctmat if par[j] else opt.gamat, dtype=np.int8)
        exp = rp.extend2(qs[j], ts[j], mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, ws[j], opt.pen_clip5, opt.zdrop, h0s[j])
        assert (got[j] == exp).all(), (j, got[j], exp)


def test_abi_symbols():
    """libbsq.so loads (no GPU needed for that) and exports everything include/bsq.h declares."""
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libbsq.so not built yet (run __graft_entry__.build())")
    hdr = open(os.path.join(os.path.dirname(capi.HERE), "include", "bsq.h")).read()
    names = set(re.findall(r"\b(bsq_[a-z0-9_]+)\s*\(", hdr))
    lib = C.CDLL(capi.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


@pytest.mark.parametrize("backend", BACKENDS)
def test_deferred_fetch_slots(ctx, backend, request):
    """The two result slots of the aligner: a batch can be fetched after the next one has run; a third run waits until a
    claimed slot has been fetched or released (bsq_aligner_result_slot / _fetch_slot / _release_slot)."""
    import threading
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    opt = bsq.default_opt()
    reads, mat, lens = _reads(ds, extra_short=False)
    n = min(len(reads), 120)
    a_in = (mat[:n // 2], lens[:n // 2], np.ones(n // 2, np.uint8))
    b_in = (mat[n // 2:n], lens[n // 2:n], np.zeros(n - n // 2, np.uint8))
    al = capi.Aligner(dx, opt)
    exp_a = al.phase1(*a_in)
    exp_b = al.phase1(*b_in)
    sa = al.stage_run(*a_in)
    sb = al.stage_run(*b_in)          # the other slot: A's regions are still on the device
    assert sa[0] != sb[0]
    done = threading.Event()

    def third():
        sc = al.stage_run(*a_in)      # would overwrite A's slot: blocks until A has been fetched
        al.release_slot(sc[0])
        done.set()

    th = threading.Thread(target=third)
    th.start()
    assert not done.wait(1.0), "the run that reuses a claimed slot did not wait"
    got_a = al.fetch_slot(*sa)
    assert done.wait(60.0), "the waiting run was not released by the fetch"
    th.join()
    got_b = al.fetch_slot(*sb)
    for got, exp in ((got_a, exp_a), (got_b, exp_b)):
        assert (got[1] == exp[1]).all() and got[0].tobytes() == exp[0].tobytes()
    al.close()
    dx.close()
