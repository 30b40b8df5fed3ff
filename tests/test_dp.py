"""Parity of the batched phase-2 dynamic programming (bsq_dp_*: SURVEY.md §8a a16 mate-rescue local alignment, a20 final
CIGAR / MD / NM / ZC / ZR) against the UNMODIFIED reference (oracle/_ref): bis_bwa_gen_cigar2 driven by the band-doubling
loop of mem_alnreg_setSAM, and ksw_align2.  Bit-exact (integer work, text)."""
import ctypes as C

import numpy as np
import pytest

import refprobe
from biscuit_b200 import capi, indexio

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]

XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _infer_bw(l1, l2, score, a, q, r):  # bwamem.h:192
    if l1 == l2 and l1 * a - score < (q + r - a) << 1:
        return 0
    w = int(float(min(l1, l2) * a - score - q) / r + 2.)
    return max(w, abs(l1 - l2))


def _ref_set_sam(rp, opt, query, job):
    """mem_alnreg_setSAM (mem_alnreg.c:40-108) around the reference's bis_bwa_gen_cigar2: final CIGAR words, MD, tags."""
    lib = rp.lib
    mat = np.array(list(opt.ctmat if job["parent"] else opt.gamat), dtype=np.int8)
    q = np.ascontiguousarray(np.minimum(query[job["qb"]:job["qe"]], 4).astype(np.uint8))
    w, last_sc = int(job["w"]), -(1 << 30)
    out = np.zeros(6, dtype=np.int32)
    cig = np.zeros(4096, dtype=np.uint32)
    md = C.create_string_buffer(16384)
    ok = 0
    for _ in range(3):
        w = min(w, opt.w << 2)
        ok = lib.refp_gen_cigar2(rp.h, _p(mat), opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, len(q), _p(q), C.c_int64(int(job["rb"])),
                                 C.c_int64(int(job["re"])), C.c_uint8(int(job["parent"])), _p(out), _p(cig), 4096, md, 16384)
        score = int(out[0])
        if score == last_sc or w == opt.w << 2 or score >= job["truesc"] - opt.a:
            break
        w <<= 1
        last_sc = score
    if not ok:
        return None
    n = int(out[1])
    words = [int(x) for x in cig[:n]]
    lead = 0
    if words and words[0] & 0xf == 2:
        lead = words[0] >> 4
        words = words[1:]
    elif words and words[-1] & 0xf == 2:
        words = words[:-1]
    if job["clip5"]:
        words = [int(job["clip5"]) << 4 | 3] + words
    if job["clip3"]:
        words = words + [int(job["clip3"]) << 4 | 3]
    return dict(cigar=words, md=md.value, NM=int(out[2]), ZC=int(out[3]), ZR=int(out[4]), bss_u=int(out[5]), score=int(out[0]), lead_del=lead)


@pytest.fixture(scope="module", params=["ds_1m", "ds_hard"])
def ctx(request):
    ds = request.getfixturevalue(request.param)
    hi = indexio.load_index(ds["fa"])
    rp = refprobe.RefProbe(ds["fa"])
    yield ds, hi, rp
    rp.close()


def _task_rows(ds, n_pairs):
    p = ds["pairs"]
    reads = [np.asarray(r, dtype=np.uint8) for r in p["r1"][:n_pairs]] + [np.asarray(r, dtype=np.uint8) for r in p["r2"][:n_pairs]]
    reads += [reads[0][:40], reads[1][:75], reads[2][20:140]]
    L = (max(len(r) for r in reads) + 15) & ~15
    mat = np.zeros((len(reads), L), dtype=np.uint8)
    for i, r in enumerate(reads):
        mat[i, :len(r)] = r
    return reads, mat, np.array([len(r) for r in reads], dtype=np.int32)


@pytest.mark.parametrize("backend", BACKENDS)
def test_cigar_jobs(ctx, backend, request):
    """Every phase-1 region of a few hundred reads becomes one CIGAR job (as mem_alnreg_setSAM would be called on it),
    plus jobs with narrowed bands, clips and a strand-bridging span."""
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    opt = bsq.default_opt()
    reads, mat, lens = _task_rows(ds, 150)
    al = capi.Aligner(dx, opt)
    rows, jobs = [], []
    for parent in (0, 1):
        regs, off = al.phase1(mat, lens, np.full(len(reads), parent))
        for t in range(len(reads)):
            for r in regs[off[t]:off[t + 1]]:
                l1, l2 = int(r["qe"] - r["qb"]), int(r["re"] - r["rb"])
                w = max(_infer_bw(l1, l2, int(r["truesc"]), opt.a, opt.o_del, opt.e_del), _infer_bw(l1, l2, int(r["truesc"]), opt.a, opt.o_ins, opt.e_ins))
                if w > opt.w:
                    w = min(w, int(r["w"]))
                is_rev = int(r["rb"]) >= hi.l_pac
                c5 = int(lens[t] - r["qe"]) if is_rev else int(r["qb"])
                c3 = int(r["qb"]) if is_rev else int(lens[t] - r["qe"])
                jobs.append((int(r["rb"]), int(r["re"]), t, int(r["qb"]), int(r["qe"]), w, int(r["truesc"]), c5, c3, int(r["parent"])))
    al.close()
    rng = np.random.default_rng(5)
    extra = []
    for j in jobs[:400:7]:  # perturbed copies: shifted ends (forces gaps at the ends), band 0, unreachable true score (band doubling)
        rb, re, t, qb, qe, w, sc, c5, c3, par = j
        d = int(rng.integers(1, 6))
        if (rb < hi.l_pac) == (re + d <= hi.l_pac) and re + d <= 2 * hi.l_pac:
            extra.append((rb, re + d, t, qb, qe, max(w, 1), sc + 40, c5, c3, par))
        if rb - d >= 0 and ((rb - d) < hi.l_pac) == (rb < hi.l_pac):
            extra.append((rb - d, re, t, qb, qe, 0, sc, 0, c3, par))
        if qe - qb > 30:
            extra.append((rb, re, t, qb + 3, qe - 5, w, sc, c5 + 3, c3 + 5, par))
    extra.append((hi.l_pac - 50, hi.l_pac + 50, 0, 0, 100, 5, 100, 0, 0, 1))  # bridges the strands: no alignment
    jobs += extra
    ja = np.zeros(len(jobs), dtype=capi.CIGAR_JOB_DTYPE)
    for k, j in enumerate(jobs):
        ja[k] = (j[0], j[1], j[2], j[3], j[4], j[5], j[6], j[7], j[8], j[9], (0, 0, 0))
    dp = capi.Dp(dx, opt)
    dp.set_reads(mat, lens)
    res, blob = dp.cigar(ja)
    c = dp.counters()
    dp.close()
    n_gapped = 0
    raw = blob.tobytes()
    for k in range(len(jobs)):
        exp = _ref_set_sam(rp, opt, reads[int(ja[k]["row"])], ja[k])
        got = res[k]
        assert got["n_cigar"] >= 0, k  # nothing here exceeds the kernel's limits
        if exp is None:
            assert got["n_cigar"] == 0, k
            continue
        n = int(got["n_cigar"])
        words = [int(x) for x in blob[int(got["off"]):int(got["off"]) + n]]
        assert words == exp["cigar"], (k, jobs[k])
        b0 = (int(got["off"]) + n) * 4
        md = raw[b0:raw.index(b"\0", b0)]
        assert md == exp["md"], (k, jobs[k], md, exp["md"])
        for f in ("NM", "ZC", "ZR", "bss_u", "score", "lead_del"):
            assert int(got[f]) == exp[f], (k, f, jobs[k])
        n_gapped += any(w & 0xf in (1, 2) for w in words)
    assert c[0] == len(jobs)
    if ds is not None and "hard" in ds["fa"]:
        assert n_gapped > 20  # the noisy set carries indels: ksw_global2 + traceback are exercised
    dx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_matesw_jobs(ctx, backend, request):
    """ksw_align2 as mate rescue calls it: reverse-complemented mate against a window of the other strand, 8-bit and
    16-bit striped kernels, with and without a hit in the window."""
    ds, hi, rp = ctx
    bsq = request.getfixturevalue(backend)
    dx = bsq.upload(hi)
    opt = bsq.default_opt()
    reads, mat, lens = _task_rows(ds, 60)
    p = ds["pairs"]
    rng = np.random.default_rng(11)
    L2 = 2 * hi.l_pac
    jobs = []
    for t in range(len(reads)):
        l_ms = int(lens[t])
        base = XSUBO | XSTART | (opt.min_seed_len * opt.a)
        for trial in range(3):
            if trial == 0:
                rb = int(rng.integers(0, L2 - 700))
            else:  # a window that holds the read's own origin on either strand: a real local hit
                cid, tpos = p["truth"][0], p["truth"][1]
                pr = t % len(tpos)
                pos = int(hi.ann_offset[int(cid[pr])] + tpos[pr])
                rb = max(0, pos - int(rng.integers(50, 300)))
                if trial == 2:
                    rb = max(hi.l_pac, L2 - rb - 600)
            re = rb + int(rng.integers(200, 650))
            if (rb < hi.l_pac) != (re <= hi.l_pac) or re > L2:
                continue
            for xb in (XBYTE if l_ms * opt.a < 250 else 0, 0):
                jobs.append((rb, re, t, base | xb, int(rng.integers(0, 2))))
    ja = np.zeros(len(jobs), dtype=capi.MATESW_JOB_DTYPE)
    for k, j in enumerate(jobs):
        ja[k] = (j[0], j[1], j[2], j[3], j[4], (0,) * 7)
    dp = capi.Dp(dx, opt)
    dp.set_reads(mat, lens)
    res = dp.matesw(ja)
    dp.close()
    out = np.zeros(7, dtype=np.int32)
    n_hit = 0
    for k, (rb, re, t, xtra, use_ga) in enumerate(jobs):
        r = reads[t]
        q = np.ascontiguousarray(np.where(r < 4, 3 - r, 4)[::-1].astype(np.uint8))
        tgt = np.zeros(re - rb + 8, dtype=np.uint8)
        n = rp.lib.refp_get_seq(rp.h, C.c_int64(rb), C.c_int64(re), _p(tgt), len(tgt))
        assert n == re - rb
        m = np.array(list(opt.gamat if use_ga else opt.ctmat), dtype=np.int8)
        rp.lib.refp_align2(len(q), _p(q), n, _p(tgt), _p(m), opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, xtra, _p(out))
        got = [int(res[k][f]) for f in ("score", "te", "qe", "score2", "te2", "tb", "qb")]
        assert res[k]["pad_"] == 0
        assert got == [int(x) for x in out], (k, jobs[k], got, out)
        n_hit += out[0] >= opt.min_seed_len and out[6] >= 0
    assert n_hit > 10
    dx.close()
