"""Pileup, vcf2bed and mergecg against the reference ITSELF: oracle/_ref/biscuit_ref_src is the unmodified
/root/reference/src/{pileup.c,bisc_utils.c,refcache.h,vcf2bed.c,mergecg.c} compiled over stand-ins for the absent htslib
and huishenlab/utils headers (oracle/ref_shim_src, oracle/Makefile).  This pins

  * the product command line (`biscuit pileup|vcf2bed|mergecg`; CUDA build under -m gpu, host side via the emulation here),
  * the CPU restatement oracle/bsq_oracle_pileup.c + bsq_oracle_vcf.c that the kernel-level tests use as their checker,

on every VCF column except QUAL / FILTER / GT / GL1 / GQ, whose arithmetic lives in utils' stats.h (not in the reference
tree; restated on both sides -> those columns are compared too, but prove nothing; `refsrc.pinned_view` is the part that does).
The committed fixture tests/golden/pileup_tiny (tools/make_golden_pileup.py) holds the reference's output for a small
case so the comparison also runs where oracle/_ref is absent."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import bamio
import oracle_plp
import refsrc
import synth
import synth_plp
from test_pileup_cli import BISCUIT, _need, _write_fasta, oracle_vcf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "pileup_tiny")

OPTION_SETS = [[], ["-s", "70000"], ["-g", "chr2:100,001-150000"], ["-N"], ["-d", "-b", "10"], ["-r", "-5", "0", "-3", "0"],
               ["-p", "-c", "-u", "-m", "0", "-a", "0"], ["-t", "3", "-n", "1", "-l", "151"], ["-E", "0.01", "-C", "0.05", "-P", "0.2", "-Q", "0.1"]]


def _case(tmp, n_bams, scale=1):
    ref_b = synth.make_reference(250_000 // scale, 1, seed=3, n_runs=2)[0][1]
    ref_a = synth.make_reference(120_000 // scale, 1, seed=8)[0][1]
    ref_c = synth.make_reference(30_000 // scale, 1, seed=9)[0][1]
    contigs = [("chr2", ref_b), ("chr10", ref_a), ("chr1", ref_c)]  # header order != name order; chr1 has no reads
    rd_b = synth_plp.make_reads(ref_b, 5000 // scale, seed=9, noise=True, n_bams=n_bams)
    rd_a = synth_plp.make_reads(ref_a, 2500 // scale, seed=10, noise=True, n_bams=n_bams)
    fa = os.path.join(tmp, "ref.fa")
    _write_fasta(fa, contigs)
    bams = []
    for s in range(n_bams):
        b = os.path.join(tmp, f"s{s}.bam")
        bamio.write_bam_from_soa(b, [(n, len(x)) for n, x in contigs], [rd_b, rd_a, None], sid=s, block=20000, tag_style=("YD", "XG", "none")[s % 3])
        bams.append(b)
    return fa, bams, contigs, {"chr2": rd_b, "chr10": rd_a}


def _compare_pileup(prog, tmp, fa, bams, opts):
    a, r = os.path.join(tmp, "a.vcf"), os.path.join(tmp, "r.vcf")
    subprocess.run([prog, "pileup", "-@", "4", "-o", a] + opts + [fa] + bams, check=True, capture_output=True)
    refsrc.run("pileup", "-@", "3", "-o", r, *opts, fa, *bams)
    ha, ba = refsrc.split_vcf(open(a, "rb").read())
    hr, br = refsrc.split_vcf(open(r, "rb").read())
    assert ha == hr, opts
    assert len(ba) == len(br) and refsrc.pinned_view(ba) == refsrc.pinned_view(br), opts
    assert ba == br, opts  # incl. the stats.h columns (same restated formulas on both sides)
    assert open(a + "_meth_average.tsv", "rb").read() == open(r + "_meth_average.tsv", "rb").read(), opts
    return len(br)


def _emu():
    import test_align_sam
    return test_align_sam.build_emu_bin()


@pytest.mark.parametrize("n_bams", [1, 3])
def test_pileup_cli_host_side_vs_reference(tmp_path, n_bams):
    """Host side (BGZF/BAM/BAI/FASTA readers, chunking, VCF text, averages) + the restatement behind the emulation,
    against the reference's pileup over nine option sets."""
    if not refsrc.available():
        pytest.skip("oracle/_ref/biscuit_ref_src not built")
    _need(oracle_plp.SO)
    fa, bams, _, _ = _case(str(tmp_path), n_bams, scale=2)
    n = [_compare_pileup(_emu(), str(tmp_path), fa, bams, o) for o in (OPTION_SETS if n_bams == 1 else OPTION_SETS[:4])]
    assert n[0] > 50000 and n[2] < n[0]


@pytest.mark.gpu
@pytest.mark.parametrize("n_bams", [1, 3])
def test_pileup_cli_vs_reference(tmp_path, n_bams):
    """`biscuit pileup` on the B200 against the reference's pileup: VCF (all option sets) and _meth_average.tsv."""
    if not refsrc.available():
        pytest.skip("oracle/_ref/biscuit_ref_src not built")
    _need(BISCUIT)
    fa, bams, _, _ = _case(str(tmp_path), n_bams)
    n = [_compare_pileup(BISCUIT, str(tmp_path), fa, bams, o) for o in OPTION_SETS]
    assert n[0] > 100000


def test_restatement_vs_reference(tmp_path):
    """The checker of the kernel-level tests (oracle_plp.region -> bsqo_plp_vcf) reproduces the reference's VCF body."""
    if not refsrc.available():
        pytest.skip("oracle/_ref/biscuit_ref_src not built")
    _need(oracle_plp.SO)
    n_bams = 2
    fa, bams, contigs, reads = _case(str(tmp_path), n_bams, scale=2)
    r = os.path.join(str(tmp_path), "r.vcf")
    refsrc.run("pileup", "-o", r, fa, *bams)
    _, body = refsrc.split_vcf(open(r, "rb").read())
    exp = []
    for name, nt4 in sorted(contigs, key=lambda c: c[0]):
        if name in reads:
            rd = dict(reads[name])
            # the BAMs were written with one strand-tag style per sample: sample 1 has XG, both carry the tag of every read
            recs = oracle_plp.region(oracle_plp.conf_default(), nt4, rd, 1, len(nt4), n_bams)
            exp.append(oracle_vcf(recs, name, n_bams)[0])
    assert b"\n".join(body) + b"\n" == b"".join(exp)


def test_vcf2bed_mergecg_vs_reference(tmp_path):
    """`biscuit vcf2bed` / `mergecg` (host C) against the reference's own, every -t target and switch."""
    if not refsrc.available():
        pytest.skip("oracle/_ref/biscuit_ref_src not built")
    _need(BISCUIT)
    tmp = str(tmp_path)
    fa, bams, _, _ = _case(tmp, 2, scale=2)
    vcf, vcfn = os.path.join(tmp, "r.vcf"), os.path.join(tmp, "rn.vcf")
    refsrc.run("pileup", "-o", vcf, fa, *bams)
    refsrc.run("pileup", "-N", "-o", vcfn, fa, *bams)
    with open(vcf, "rb") as fi, gzip.open(vcf + ".gz", "wb") as fo:
        fo.write(fi.read())
    cases = [(vcf, a.split()) for a in ("-t cg", "-t ch -k 3", "-t c", "-t hcg", "-t gch", "-t snp", "-t cg -e", "-t cg -c", "-t c -s ALL", "-t cg -s LAST -k 5",
                                       "-t snp -s ALL", "-t cg -s s1,s0 -e -c", "-t CG -k 2")]
    cases += [(vcfn, ["-t", "hcg"]), (vcfn, ["-t", "gch", "-e"]), (vcf + ".gz", ["-t", "cg"])]
    beds = {}
    for f, args in cases:
        exp = refsrc.run("vcf2bed", *args, f).stdout
        got = subprocess.run([BISCUIT, "vcf2bed", *args, f], capture_output=True, check=True).stdout
        assert got == exp, args
        beds[(os.path.basename(f), " ".join(args))] = exp
    assert beds[("r.vcf", "-t cg")].count(b"\n") > 3000 and beds[("rn.vcf", "-t hcg")].count(b"\n") > 3000
    for key, name in ((("r.vcf", "-t cg"), "cg1"), (("r.vcf", "-t c -s ALL"), "c2"), (("rn.vcf", "-t hcg"), "hcg"), (("r.vcf", "-t cg -s s1,s0 -e -c"), "ctx")):
        open(os.path.join(tmp, name + ".bed"), "wb").write(beds[key])
    for args in (["cg1.bed"], ["-c", "cg1.bed"], ["-k", "4", "c2.bed"], ["-N", "hcg.bed"], ["-N", "-c", "-k", "2", "c2.bed"]):
        args = args[:-1] + [fa, os.path.join(tmp, args[-1])]
        exp = refsrc.run("mergecg", *args).stdout
        got = subprocess.run([BISCUIT, "mergecg", *args], capture_output=True, check=True).stdout
        assert got == exp and exp.count(b"\n") > 500, args


def _golden_compare(prog, tmp):
    fa = os.path.join(tmp, "ref.fa")
    with gzip.open(os.path.join(GOLD, "ref.fa.gz"), "rb") as fi, open(fa, "wb") as fo:
        fo.write(fi.read())
    bams = [os.path.join(GOLD, "s0.bam"), os.path.join(GOLD, "s1.bam")]
    for tag, opts in (("default", []), ("nome_step", ["-N", "-s", "3000"])):
        out = os.path.join(tmp, tag + ".vcf")
        subprocess.run([prog, "pileup", "-@", "2", "-o", out] + opts + [fa] + bams, check=True, capture_output=True)
        h, b = refsrc.split_vcf(open(out, "rb").read())
        eh, eb = refsrc.split_vcf(gzip.open(os.path.join(GOLD, tag + ".vcf.gz"), "rb").read())
        # the sample columns of #CHROM come from the BAM paths' basenames: same here; ##reference names the FASTA path
        h = [l for l in h if not l.startswith(b"##reference")]
        eh = [l for l in eh if not l.startswith(b"##reference")]
        assert h == eh and b == eb, tag
        tsv = [l.split("\t", 1)[1] for l in open(out + "_meth_average.tsv").read().splitlines()]
        etsv = [l.split("\t", 1)[1] for l in gzip.open(os.path.join(GOLD, tag + ".tsv.gz"), "rt").read().splitlines()]
        assert tsv == etsv, tag
        if tag == "default":
            bed = subprocess.run([prog, "vcf2bed", "-t", "cg", "-s", "ALL", out], capture_output=True, check=True).stdout
            assert bed == gzip.open(os.path.join(GOLD, "cg.bed.gz"), "rb").read()
            bedf = os.path.join(tmp, "cg.bed")
            open(bedf, "wb").write(bed)
            mg = subprocess.run([prog, "mergecg", fa, bedf], capture_output=True, check=True).stdout
            assert mg == gzip.open(os.path.join(GOLD, "cg_merged.bed.gz"), "rb").read()


def test_golden_fixture_host_side(tmp_path):
    """Committed reference output (tests/golden/pileup_tiny) vs the host programs over the emulation."""
    _need(oracle_plp.SO, os.path.join(GOLD, "default.vcf.gz"))
    _golden_compare(_emu(), str(tmp_path))


@pytest.mark.gpu
def test_golden_fixture_gpu(tmp_path):
    """Committed reference output (tests/golden/pileup_tiny) vs `biscuit pileup | vcf2bed | mergecg` on the B200."""
    _need(BISCUIT, os.path.join(GOLD, "default.vcf.gz"))
    _golden_compare(BISCUIT, str(tmp_path))


def _hard_clip_reads():
    from test_pileup import _mini_reads
    ref = np.tile(np.array([1, 2, 0, 3, 1, 1, 2, 2], np.uint8), 40)  # CGATCCGG...
    s = "CGATCCGG" * 6
    e = [dict(pos=8, seq=s, bss=0, cigar=[(16, 5), (48, 0)]),            # leading H: bases shifted by 16, last 16 events past SEQ
         dict(pos=16, seq=s, bss=1, cigar=[(48, 0), (30, 5)]),           # trailing H: no effect
         dict(pos=24, seq=s[:40], bss=0, cigar=[(8, 5), (20, 0), (3, 1), (2, 2), (17, 0)]),
         dict(pos=32, seq=s, bss=0, cigar=[(48, 0)]),
         dict(pos=40, seq="", bss=0, cigar=[(48, 0)]),                   # SEQ '*' with a CIGAR: fails min_read_len, touches nothing
         dict(pos=48, seq=s, bss=-1, cigar=[(12, 5), (48, 0)]),          # strand inferred: bases past SEQ are skipped
         dict(pos=56, seq=s, bss=0, cigar=[(48, 0)])]
    return ref, _mini_reads(e, 48)


def test_hard_clip_restatement():
    """Reference semantics of H (qpos advances, pileup.c:822-824) restated without the out-of-range reads."""
    _need(oracle_plp.SO)
    ref, rd = _hard_clip_reads()
    c = oracle_plp.conf_default()
    c.min_dist_end_5p = c.min_dist_end_3p = 0
    # read 0 alone covers 9..56 (1-based) but only its first 32 events have a base behind them: nothing is emitted past 40
    only0 = oracle_plp.region(c, ref, {**rd, "n_reads": 1}, 1, len(ref))
    assert len(only0) > 5 and max(int(r["pos"]) for r in only0) <= 9 + 32 - 1
    # ... yet the 16 events past SEQ count towards DP (pileup.c:572): seen next to read 3 (33..80), the only other read there
    sub = {**rd, "n_reads": 4}
    two = {int(r["pos"]): r for r in oracle_plp.region(c, ref, sub, 1, len(ref))}
    hit = [p for p in range(41, 57) if p in two]
    assert hit and all(int(two[p]["dp"]) - int(two[p]["base"].sum()) >= 1 for p in hit)
    assert all(int(two[p]["dp"]) - int(two[p]["base"].sum()) == 1 for p in hit if p <= 44)  # read 2's own overrun starts later


@pytest.mark.gpu
def test_hard_clip_gpu(cuda):
    from biscuit_b200 import plp
    ref, rd = _hard_clip_reads()
    pl = plp.Pileup(cuda, 1)
    conf = pl.default_conf()
    pl.set_contig(ref)
    for d5 in (3, 0):
        conf.min_dist_end_5p = conf.min_dist_end_3p = d5
        got = pl.region(conf, rd, 1, len(ref))
        exp = oracle_plp.region(conf, ref, rd, 1, len(ref), 1)
        assert len(exp) > 20 and got.tobytes() == exp.tobytes()
    pl.close()


def _run_ranks(prog, tmp, fa, bams, world, opts=(), devices=None):
    out = os.path.join(tmp, f"w{world}.vcf")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(devices[r] if devices else 0), MASTER_PORT="29512")
        procs.append(subprocess.Popen([prog, "pileup", "-@", "2", "-o", out, *opts, fa, *bams], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    for p in procs:
        _, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-2000:]
    left = [f for f in os.listdir(tmp) if ".part" in f or ".done" in f or ".stats" in f]
    assert not left, left
    return out


def _check_sharded(prog, tmp, devices=None):
    fa, bams, _, _ = _case(tmp, 2, scale=2)
    one = os.path.join(tmp, "one.vcf")
    subprocess.run([prog, "pileup", "-@", "2", "-o", one, fa] + bams, check=True, capture_output=True)
    for world in (2, 3):
        if devices and world > len(devices):
            continue
        out = _run_ranks(prog, tmp, fa, bams, world, devices=devices)
        assert refsrc.split_vcf(open(out, "rb").read()) == refsrc.split_vcf(open(one, "rb").read())
        assert open(out + "_meth_average.tsv", "rb").read() == open(one + "_meth_average.tsv", "rb").read()
    # a region is not sharded: rank 0 does it, the result is the one-process result
    reg = ["-g", "chr2:10,001-60000"]
    subprocess.run([prog, "pileup", "-@", "2", "-o", one] + reg + [fa] + bams, check=True, capture_output=True)
    out = _run_ranks(prog, tmp, fa, bams, 2, opts=reg, devices=devices)
    assert refsrc.split_vcf(open(out, "rb").read())[1] == refsrc.split_vcf(open(one, "rb").read())[1]


def test_pileup_ranks_host_side(tmp_path):
    """`biscuit pileup` as one process per rank (RANK / WORLD_SIZE / LOCAL_RANK): contigs dealt round-robin, rank 0 merges the
    VCF bodies in contig-name order and adds the per-contig statistics -- same files as a one-process run."""
    _need(oracle_plp.SO)
    _check_sharded(_emu(), str(tmp_path))


@pytest.mark.gpu
def test_pileup_two_gpus(tmp_path):
    import torch
    _need(BISCUIT)
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _check_sharded(BISCUIT, str(tmp_path), devices=[0, 1])
