"""End-to-end `biscuit align` parity: SAM byte-identical to the unmodified reference (oracle/_ref/biscuit_ref)
modulo the @PG line, over the reference's own command-line boundary (B3).

* not-gpu: host code (biscuit_b200/host/*.c) linked against the test-only host emulation of libbsq
* gpu:     the product binary biscuit_b200/host/biscuit (links libbsq.so, CUDA)
"""
import os
import subprocess

import numpy as np
import pytest

import refprobe
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "biscuit_b200", "host")
EMU_BIN = os.path.join(ROOT, "tests", "hostemu", "biscuit_hostemu")
GPU_BIN = os.path.join(HOST, "biscuit")


def build_emu_bin():
    import conftest
    conftest.build_hostemu()
    srcs = [os.path.join(HOST, f) for f in ("bq_core.c", "bq_phase2.c", "bq_pipe.c", "bq_io.c", "bq_main.c", "bq_bam.c", "bq_pileup.c", "bq_vcf2bed.c", "bq_sortbam.c")]
    deps = srcs + [os.path.join(HOST, "bq.h"), os.path.join(HOST, "bq_sort.h"), os.path.join(HOST, "bq_plp.h"), os.path.join(ROOT, "tests", "hostemu", "libbsq_hostemu.so")]
    if not os.path.exists(EMU_BIN) or any(os.path.getmtime(d) > os.path.getmtime(EMU_BIN) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-g", "-std=gnu11", "-o", EMU_BIN] + srcs +
                              ["-L" + os.path.dirname(EMU_BIN), "-lbsq_hostemu", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread", "-lm"])
    return EMU_BIN


@pytest.fixture(scope="module")
def hard_set(tmp_path_factory):
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path_factory.mktemp("sam"))
    ref = synth.make_reference(300_037, 5, seed=9, n_runs=3)
    fa = os.path.join(d, "hard.fa")
    synth.write_fasta(fa, ref)
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = synth.simulate_pairs(ref, 3000, seed=3, sub_rate=0.03, indel_rate=0.006, n_rate=0.002, qual="mixed")
    rng = np.random.default_rng(1)
    r1, r2 = p["r1"].copy(), p["r2"].copy()
    k = rng.random(len(r2)) < 0.10  # unrelated mates: mate rescue, unpaired output
    r2[k] = rng.integers(0, 4, size=(int(k.sum()), 150))
    k = rng.random(len(r2)) < 0.05
    r1[k] = rng.integers(0, 4, size=(int(k.sum()), 150))
    r1[:20, :] = 0  # low-complexity reads
    r2[20:40, 30:120] = 3
    f1, f2 = os.path.join(d, "h1.fq"), os.path.join(d, "h2.fq")
    synth.write_fastq(f1, r1, p["q1"], suffix="/1")
    synth.write_fastq(f2, r2, p["q2"], suffix="/2")
    return fa, f1, f2


def _sam(binary, args, env=None):
    out = subprocess.run([binary, "align"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True,
                         env=dict(os.environ, **env) if env else None).stdout
    return b"\n".join(ln for ln in out.split(b"\n") if not ln.startswith(b"@PG"))


CASES = [[], ["-b", "1"], ["-a"], ["-5", "4", "-3", "7", "-z", "15"], ["-A", "2"], ["-I", "450,40"], ["-S"], ["-P"], ["-M", "-Y"],
         ["-R", "@RG\\tID:x\\tSM:y"]]


@pytest.mark.parametrize("extra", CASES, ids=[" ".join(c) or "default" for c in CASES])
def test_sam_identical_hostemu(hard_set, extra):
    fa, f1, f2 = hard_set
    args = ["-@", "4"] + extra + [fa, f1, f2]
    assert _sam(build_emu_bin(), args) == _sam(refprobe.REF_BIN, args)


def test_sam_many_small_batches_hostemu(hard_set):
    """Eleven batches instead of one (BQ_CHUNK_SIZE test hook): with the insert-size distribution given (-I) nothing in
    the output depends on where batches end, so the text must still equal the reference's single-batch run.  Exercises
    the buffers that are recycled from batch to batch (region pool, SAM slabs, read slabs) and the parked workers."""
    fa, f1, f2 = hard_set
    args = ["-@", "4", "-I", "450,40", fa, f1, f2]
    assert _sam(build_emu_bin(), args, env={"BQ_CHUNK_SIZE": "20000"}) == _sam(refprobe.REF_BIN, args)


def test_fastq_shapes_hostemu(hard_set, tmp_path):
    """FASTQ records the fast path of the reader must leave to the general kseq path -- comments after the name,
    CR/LF line ends, sequences and qualities folded over several lines, FASTA records without qualities, a last
    record without a newline -- mixed with ordinary ones: same SAM as the reference, and the same with the fast path
    switched off (BQ_FQ_SLOW)."""
    fa, f1, f2 = hard_set
    out = []
    for k, fn in enumerate((f1, f2)):
        lines = open(fn).read().split("\n")
        recs = [lines[4 * i:4 * i + 4] for i in range(400)]
        txt = []
        for i, (nm, sq, pl, ql) in enumerate(recs):
            if i % 7 == 1:
                nm += " a comment"
            if i % 11 == 2:
                txt.append("\r\n".join((nm, sq, pl, ql)) + "\r\n")
            elif i % 13 == 3:
                txt.append("\n".join((nm, sq[:70], sq[70:], "+" + nm[1:], ql[:40], ql[40:])) + "\n")
            elif i % 17 == 4:
                txt.append(">" + nm[1:] + "\n" + sq + "\n")
            else:
                txt.append("\n".join((nm, sq, pl, ql)) + "\n")
        body = "".join(txt)
        o = str(tmp_path / f"odd{k}.fq")
        open(o, "w").write(body[:-1])  # no newline after the last record
        out.append(o)
    args = ["-@", "2", fa] + out
    mine = _sam(build_emu_bin(), args)
    assert mine == _sam(refprobe.REF_BIN, args)
    assert mine == _sam(build_emu_bin(), args, env={"BQ_FQ_SLOW": "1"})
    assert mine.count(b"\n") > 800


def test_gzip_pair_files_hostemu(hard_set, tmp_path):
    """gzip-compressed pair files (kseq over gzread in the reference): the second file is inflated and parsed by the
    reader's helper thread, the batches are read two ahead of the pipeline -- same SAM as the reference, with the helper
    and without it."""
    import gzip
    import shutil
    fa, f1, f2 = hard_set
    gz = []
    for k, fn in enumerate((f1, f2)):
        o = str(tmp_path / f"p{k}.fq.gz")
        with open(fn, "rb") as fi, gzip.open(o, "wb", compresslevel=1) as fo:
            shutil.copyfileobj(fi, fo)
        gz.append(o)
    args = ["-@", "3", fa] + gz
    ref = _sam(refprobe.REF_BIN, args)
    assert _sam(build_emu_bin(), args) == ref
    assert _sam(build_emu_bin(), args, env={"BQ_FQ_NO_AHEAD": "1"}) == ref


def test_sam_identical_single_end_hostemu(hard_set):
    fa, f1, _ = hard_set
    args = ["-@", "4", fa, f1]
    assert _sam(build_emu_bin(), args) == _sam(refprobe.REF_BIN, args)


def _interleave(f1, f2, out):
    """Interleaved FASTQ for -p: most pairs complete, some reads without their mate (single-end)."""
    a, b = open(f1).read().split("\n"), open(f2).read().split("\n")
    with open(out, "w") as fo:
        for i in range(len(a) // 4):
            ra, rb = a[4 * i:4 * i + 4], b[4 * i:4 * i + 4]
            if i % 7 != 3:
                fo.write("\n".join(ra) + "\n")
            if i % 11 != 5:
                fo.write("\n".join(rb) + "\n")
    return out


@pytest.mark.parametrize("extra", [[], ["-K", "70000"], ["-I", "450,40"]], ids=["default", "adaptor -K", "-I"])
def test_sam_identical_smart_pairing_hostemu(hard_set, tmp_path, extra):
    """-p (MEM_F_SMARTPE, align.c:109-143): interleaved input, pairs by equal neighbouring names, the rest single-end."""
    fa, f1, f2 = hard_set
    fq = _interleave(f1, f2, str(tmp_path / "il.fq"))
    args = ["-@", "4", "-p"] + extra + [fa, fq]
    mine = _sam(build_emu_bin(), args)
    assert mine == _sam(refprobe.REF_BIN, args)
    assert mine.count(b"\n") > 5000


@pytest.mark.gpu
def test_sam_identical_smart_pairing_gpu(hard_set, tmp_path):
    fa, f1, f2 = hard_set
    fq = _interleave(f1, f2, str(tmp_path / "il.fq"))
    args = ["-@", "4", "-p", "-K", "200000", fa, fq]
    assert _sam(GPU_BIN, args) == _sam(refprobe.REF_BIN, args)


def test_sam_clean_1m_hostemu(ds_1m, tmp_path):
    """BASELINE.json configs[0] shape: clean 2x150 pairs vs a 1 Mb reference."""
    p = ds_1m["pairs"]
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, p["r1"], p["q1"])
    synth.write_fastq(f2, p["r2"], p["q2"])
    args = ["-@", "4", ds_1m["fa"], f1, f2]
    assert _sam(build_emu_bin(), args) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", CASES, ids=[" ".join(c) or "default" for c in CASES])
def test_sam_identical_gpu(hard_set, extra):
    """The same ten option sets as the host-side suite, through the CUDA build."""
    if not os.path.exists(GPU_BIN):
        pytest.fail("biscuit_b200/host/biscuit not built: run __graft_entry__.build()")
    fa, f1, f2 = hard_set
    args = ["-@", "4"] + extra + [fa, f1, f2]
    assert _sam(GPU_BIN, args) == _sam(refprobe.REF_BIN, args)


def _truncated_pair_files(hard_set, tmp_path, which):
    """Copies of the two FASTQ files with the last 37 records of one of them cut off."""
    fa, f1, f2 = hard_set
    out = []
    for k, f in enumerate((f1, f2)):
        lines = open(f).read().split("\n")
        if k == which:
            lines = lines[:len(lines) - 1 - 4 * 37] + [""]
        o = str(tmp_path / f"t{k}.fq")
        open(o, "w").write("\n".join(lines))
        out.append(o)
    return fa, out[0], out[1]


@pytest.mark.parametrize("which", [0, 1], ids=["first shorter", "second shorter"])
def test_unequal_pair_files_hostemu(hard_set, tmp_path, which):
    """One file of the pair ends early (bis_bseq_read, bwa.c:831-846: warn, align what is paired): same records as the
    reference, with the second file parsed ahead by the helper thread and with it switched off."""
    fa, f1, f2 = _truncated_pair_files(hard_set, tmp_path, which)
    args = ["-@", "3", fa, f1, f2]
    ref = _sam(refprobe.REF_BIN, args)
    assert _sam(build_emu_bin(), args) == ref
    assert _sam(build_emu_bin(), args, env={"BQ_FQ_NO_AHEAD": "1"}) == ref
    # many small batches: the helper runs across batch boundaries; same text with and without it
    assert _sam(build_emu_bin(), args, env={"BQ_CHUNK_SIZE": "9000"}) == _sam(build_emu_bin(), args, env={"BQ_CHUNK_SIZE": "9000", "BQ_FQ_NO_AHEAD": "1"})


def _dp_stats(binary, args):
    """Counters of the batched phase-2 DP as `biscuit align` reports them under BQ_TIMING."""
    import re
    err = subprocess.run([binary, "align"] + args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True,
                         env=dict(os.environ, BQ_TIMING="1")).stderr.decode()
    m = re.search(r"phase-2 DP on the GPU: (\d+) CIGAR jobs, (\d+) used, (\d+) setSAM calls on the host; mate rescue: (\d+) jobs, (\d+) used, (\d+) on the host", err)
    assert m, err[-2000:]
    return [int(x) for x in m.groups()]


def _check_dp_used(binary, hard_set):
    """The final CIGARs and the mate-rescue alignments come from the bsq_dp_* calls, not from the host routines that
    remain for regions the prediction misses (noisy set: 10 % unrelated mates, indels)."""
    fa, f1, f2 = hard_set
    cig_jobs, cig_used, cig_host, ms_jobs, ms_used, ms_host = _dp_stats(binary, ["-@", "4", fa, f1, f2])
    assert cig_used > 5000 and cig_jobs >= cig_used
    assert cig_host <= 0.02 * cig_used, (cig_used, cig_host)
    assert ms_used > 100 and ms_jobs >= ms_used
    assert ms_host <= 0.02 * ms_used + 5, (ms_used, ms_host)


def test_phase2_dp_is_used_hostemu(hard_set):
    _check_dp_used(build_emu_bin(), hard_set)


@pytest.mark.gpu
def test_phase2_dp_is_used_gpu(hard_set):
    _check_dp_used(GPU_BIN, hard_set)


@pytest.mark.gpu
def test_sam_identical_single_end_gpu(hard_set):
    fa, f1, _ = hard_set
    args = ["-@", "4", fa, f1]
    assert _sam(GPU_BIN, args) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
def test_sam_many_small_batches_gpu(hard_set):
    """Eleven batches through the CUDA build (device buffers, page-locked slots and SAM slabs recycled batch to batch)."""
    fa, f1, f2 = hard_set
    args = ["-@", "4", "-I", "450,40", fa, f1, f2]
    assert _sam(GPU_BIN, args, env={"BQ_CHUNK_SIZE": "20000"}) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
def test_index_cli_gpu(hard_set, tmp_path):
    """`biscuit index` (GPU suffix sorting) writes the same seven files as the reference, byte for byte."""
    fa, _, _ = hard_set
    mine = str(tmp_path / "mine.fa")
    subprocess.check_call(["cp", fa, mine])
    subprocess.check_call([GPU_BIN, "index", mine], stderr=subprocess.DEVNULL)
    for ext in (".par.bwt", ".dau.bwt", ".par.sa", ".dau.sa", ".bis.pac", ".bis.ann", ".bis.amb"):
        assert open(mine + ext, "rb").read() == open(fa + ext, "rb").read(), ext


@pytest.fixture(scope="module")
def repeat_set(tmp_path_factory):
    """Tandem and dispersed repeats: many equally good hits (XA/XB tags, mapQ 0, secondary marking), mate rescue
    among copies, SMEM intervals above max_occ."""
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path_factory.mktemp("samrep"))
    rng = np.random.default_rng(4)
    unit = rng.integers(0, 4, size=400, dtype=np.uint8)
    tandem = np.tile(unit, 30)
    bg = rng.integers(0, 4, size=150_000, dtype=np.uint8)
    elem = rng.integers(0, 4, size=600, dtype=np.uint8)
    for k in range(12):  # a dispersed element with small differences between copies
        p0 = 5000 + k * 11000
        cp = elem.copy()
        m = rng.random(len(cp)) < 0.01
        cp[m] = (cp[m] + 1) % 4
        bg[p0:p0 + len(cp)] = cp
    ref = [("chrBg", bg), ("chrTandem", np.concatenate([rng.integers(0, 4, size=3000, dtype=np.uint8), tandem,
                                                        rng.integers(0, 4, size=3000, dtype=np.uint8)]).astype(np.uint8))]
    fa = os.path.join(d, "rep.fa")
    synth.write_fasta(fa, ref)
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = synth.simulate_pairs(ref, 1500, seed=8, sub_rate=0.01, indel_rate=0.002, qual="mixed")
    f1, f2 = os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")
    synth.write_fastq(f1, p["r1"], p["q1"], suffix="/1")
    synth.write_fastq(f2, p["r2"], p["q2"], suffix="/2")
    return fa, f1, f2


@pytest.mark.parametrize("extra", [[], ["-a"], ["-K", "60000"]], ids=["default", "-a", "adaptor -K"])
def test_sam_identical_repeats_hostemu(repeat_set, extra):
    fa, f1, f2 = repeat_set
    args = ["-@", "3"] + extra + [fa, f1, f2]
    a, b = _sam(build_emu_bin(), args), _sam(refprobe.REF_BIN, args)
    assert a == b
    if "-a" not in extra:  # with -a the alternative hits are separate records instead of XA tags
        assert b"XA:Z:" in b and b"XB:Z:" in b  # the data set really has multi-hit reads
    else:
        assert sum(1 for ln in b.split(b"\n") if ln and not ln.startswith(b"@") and int(ln.split(b"\t")[1]) & 0x100) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["-a"], ["-K", "60000"]], ids=["default", "-a", "adaptor -K"])
def test_sam_identical_repeats_gpu(repeat_set, extra):
    fa, f1, f2 = repeat_set
    args = ["-@", "3"] + extra + [fa, f1, f2]
    assert _sam(GPU_BIN, args) == _sam(refprobe.REF_BIN, args)


def test_long_read_and_name_mismatch_hostemu(hard_set, tmp_path):
    """A read beyond the device kernels' length limit does not fail the batch: it is reported unaligned with a warning and
    every other pair is aligned as usual.  Desynchronised FASTQ files stop the run with the reference's message
    (check_paired_read_names, lib/aln/bwamem.c:210-216)."""
    fa, f1, f2 = hard_set
    a, b = open(f1).read().split("\n"), open(f2).read().split("\n")
    n = 200
    a, b = a[:4 * n], b[:4 * n]
    long_seq = (a[4 * 7 + 1] * 3)[:300]
    a[4 * 7 + 1], a[4 * 7 + 3] = long_seq, "I" * 300
    g1, g2 = str(tmp_path / "l1.fq"), str(tmp_path / "l2.fq")
    open(g1, "w").write("\n".join(a) + "\n")
    open(g2, "w").write("\n".join(b) + "\n")
    r = subprocess.run([build_emu_bin(), "align", "-@", "2", fa, g1, g2], capture_output=True)
    assert r.returncode == 0 and b"longer than 256 bases" in r.stderr
    recs = [ln.split(b"\t") for ln in r.stdout.split(b"\n") if ln and not ln.startswith(b"@")]
    name = a[4 * 7][1:].split("/")[0].encode()
    mine = [f for f in recs if f[0] == name and int(f[1]) & 0x40]
    assert len(mine) == 1 and int(mine[0][1]) & 0x4 and len(mine[0][9]) == 300
    # all other pairs: same records as a reference run in which that read is an unalignable one of ordinary length (the
    # pair keeps its place in the batch, so the read ids behind the hash tie-breaks are the same)
    h1 = str(tmp_path / "k1.fq")
    c = list(a)
    c[4 * 7 + 1], c[4 * 7 + 3] = "ACGT" * 5 + "N" * 110 + "TGCA" * 5, "I" * 150
    open(h1, "w").write("\n".join(c) + "\n")
    ref = subprocess.run([refprobe.REF_BIN, "align", "-@", "2", "-I", "450,40", fa, h1, g2], capture_output=True, check=True).stdout
    got = subprocess.run([build_emu_bin(), "align", "-@", "2", "-I", "450,40", fa, g1, g2], capture_output=True, check=True).stdout
    strip = lambda out: [ln for ln in out.split(b"\n") if ln and not ln.startswith(b"@") and ln.split(b"\t")[0] != name]  # noqa: E731
    assert strip(got) == strip(ref)
    # names that do not match
    b[4 * 3] = "@somebody_else/2"
    open(g2, "w").write("\n".join(b) + "\n")
    r = subprocess.run([build_emu_bin(), "align", "-@", "2", fa, f1, g2], capture_output=True)
    assert r.returncode != 0 and b"paired reads have different names" in r.stderr


def test_sam_identical_three_lanes_hostemu(hard_set):
    """`-G 0,0,0`: three aligner contexts, batches dealt round-robin (the multi-GPU pipeline of bq_pipe.c, here over the
    emulation).  Eleven small batches; with the insert-size distribution given nothing depends on the batch borders."""
    fa, f1, f2 = hard_set
    args = ["-@", "4", "-I", "450,40", fa, f1, f2]
    assert _sam(build_emu_bin(), ["-G", "0,0,0"] + args, env={"BQ_CHUNK_SIZE": "20000"}) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
def test_sam_identical_two_gpus(hard_set):
    """`biscuit align -G 0-1`: FASTQ batches dealt to two GPUs, output identical to the one-GPU run and to the reference."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    fa, f1, f2 = hard_set
    args = ["-@", "4", fa, f1, f2]
    two = _sam(GPU_BIN, ["-G", "0-1"] + args, env={"BQ_CHUNK_SIZE": "20000"})
    assert two == _sam(GPU_BIN, args, env={"BQ_CHUNK_SIZE": "20000"})
    args = ["-@", "4", "-I", "450,40", fa, f1, f2]
    assert _sam(GPU_BIN, ["-G", "0,1"] + args, env={"BQ_CHUNK_SIZE": "20000"}) == _sam(refprobe.REF_BIN, args)
