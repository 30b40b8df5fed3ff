"""`biscuit pileup | vcf2bed | mergecg`: host side (BGZF/BAM/BAI/FASTA readers, VCF text, methylation averages)
and, on the GPU box, the whole command line against the oracle pipeline (bsqo_plp_region -> bsqo_plp_vcf).

CPU tests never run a pileup: they check the readers against files written by tools/bamio.py, and the product
formatter (bq_plp_format) against the oracle's text restatement on records made by the oracle.  The genotype
columns (QUAL FILTER GT GL1 GQ) are "parity unpinned" (absent huishenlab/utils stats.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import bamio
import oracle_plp
import synth
import synth_plp
from biscuit_b200 import plp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BISCUIT = os.path.join(ROOT, "biscuit_b200", "host", "biscuit")
HOSTLIB = os.path.join(ROOT, "biscuit_b200", "host", "libbiscuit_host.so")
I32MIN = np.iinfo(np.int32).min


class VcfConf(C.Structure):
    _fields_ = [("error", C.c_double), ("contam", C.c_double), ("prior0", C.c_double), ("prior1", C.c_double), ("prior2", C.c_double),
                ("is_nome", C.c_int32), ("pad_", C.c_int32)]


class FmtConf(C.Structure):
    _fields_ = [("n_bams", C.c_int), ("is_nome", C.c_int), ("n_threads", C.c_int), ("error", C.c_double), ("contam", C.c_double),
                ("prior0", C.c_double), ("prior1", C.c_double), ("prior2", C.c_double)]


class Str(C.Structure):
    _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]


def oracle_vcf(recs, chrm, n_bams, is_nome=0, w0=1, step=100000):
    """Text + per-window statistics through the oracle, window by window like process_func/write_func."""
    lib = C.CDLL(oracle_plp.SO)
    lib.bsqo_plp_vcf.restype = C.c_void_p
    cf = VcfConf(0.001, 0.01, 1.0 - 0.33333 - 0.33333, 0.33333, 0.33333, is_nome, 0)
    n_loci = len(recs) // n_bams
    pos = recs["pos"][::n_bams]
    text = []
    beta_tot, cnt_tot = np.zeros(n_bams * 6), np.zeros(n_bams * 6, np.int64)
    lo = 0
    while lo < n_loci:
        w = (int(pos[lo]) - w0) // step
        hi = lo
        while hi < n_loci and (int(pos[hi]) - w0) // step == w:
            hi += 1
        b, c = np.zeros(n_bams * 6), np.zeros(n_bams * 6, np.int64)
        sub = np.ascontiguousarray(recs[lo * n_bams:hi * n_bams])
        p = lib.bsqo_plp_vcf(C.byref(cf), chrm.encode(), sub.ctypes.data_as(C.c_void_p), C.c_int64(hi - lo), C.c_int(n_bams),
                             b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
        text.append(C.string_at(p))
        lib.bsqo_free(C.c_void_p(p))
        beta_tot += b
        cnt_tot += c
        lo = hi
    return b"".join(text), beta_tot, cnt_tot


def product_format(recs, chrm, n_bams, n_threads, is_nome=0, w0=1, step=100000):
    lib = C.CDLL(HOSTLIB)
    n_loci = len(recs) // n_bams
    n_win = (int(recs["pos"].max()) - w0) // step + 1 if n_loci else 1
    cf = FmtConf(n_bams, is_nome, n_threads, 0.001, 0.01, 1.0 - 0.33333 - 0.33333, 0.33333, 0.33333)
    out = Str(0, 0, None)
    wb, wc = np.zeros((n_win, n_bams * 6)), np.zeros((n_win, n_bams * 6), np.int64)
    r = np.ascontiguousarray(recs)
    lib.bq_plp_format(C.byref(cf), chrm.encode(), r.ctypes.data_as(C.c_void_p), C.c_int64(n_loci), C.c_int64(w0), C.c_int64(step), C.c_int(n_win),
                      C.byref(out), wb.ctypes.data_as(C.c_void_p), wc.ctypes.data_as(C.c_void_p))
    text = C.string_at(out.s, out.l) if out.s else b""
    beta = np.zeros(n_bams * 6)
    for w in range(n_win):  # block order, as write_func adds the records
        beta += wb[w]
    return text, beta, wc.sum(axis=0)


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip(f"{os.path.relpath(p, ROOT)} not built")


@pytest.fixture(scope="module")
def plp_case():
    ref = synth.make_reference(300_000, 1, seed=3, n_runs=3)[0][1]
    rd = synth_plp.make_reads(ref, 6000, seed=9, noise=True, n_bams=2)
    return ref, rd


def test_format_matches_oracle_text(plp_case):
    _need(oracle_plp.SO, HOSTLIB)
    ref, rd = plp_case
    for n_bams in (1, 2):
        rd1 = dict(rd)
        if n_bams == 1:
            rd1["sid"] = np.zeros_like(rd["sid"])
        recs = oracle_plp.region(oracle_plp.conf_default(), ref, rd1, 1, len(ref), n_bams)
        assert len(recs) > 1000
        exp, eb, ec = oracle_vcf(recs, "chrT", n_bams, step=50000)
        for nt in (1, 5):
            got, gb, gc = product_format(recs, "chrT", n_bams, nt, step=50000)
            assert got == exp
            assert gb.tobytes() == eb.tobytes() and gc.tolist() == ec.tolist()
        # spot-check the pinned columns of one CpG line by hand
        line = exp.split(b"\n")[0].split(b"\t")
        assert line[0] == b"chrT" and line[2] == b"." and line[8].startswith(b"GT:GL1:GQ:DP:SP")


def test_format_nome_and_large_counts():
    _need(oracle_plp.SO, HOSTLIB)
    recs = np.zeros(3, plp.REC_DTYPE)
    # a deep CpG (counts beyond the cached tables), a SNP with ambiguous alt, a G in NOMe mode
    recs[0] = (100, 900, (700, 150, 50), (0, 700, 0, 0, 0, 150, 0), (0, 850, 0, 0, 0, 0, 0), 1, -1, 0, 1, b"AACGT", 1, (0, 0))
    recs[1] = (101, 40, (0, 0, 40), (30, 0, 0, 0, 0, 10, 0), (30, 0, 0, 0, 0, 10, 0), 0, 5, 6, 0, b"NNNNN", 0, (0, 0))
    recs[2] = (200, 300, (1, 299, 0), (299, 0, 1, 0, 0, 0, 0), (299, 0, 1, 0, 0, 0, 0), 2, 0, 3, 0, b"AGCGT", 0, (0, 0))
    for nome in (0, 1):
        exp, eb, ec = oracle_vcf(recs, "c", 1, is_nome=nome)
        got, gb, gc = product_format(recs, "c", 1, 2, is_nome=nome)
        assert got == exp and gb.tobytes() == eb.tobytes() and gc.tolist() == ec.tolist()
    assert b"CX=GCG" in oracle_vcf(recs, "c", 1, is_nome=1)[0] and b";AB=Y" in exp and b":850:0.824" in exp


def _write_fasta(path, contigs):
    with open(path, "w") as fh:
        for name, nt4 in contigs:
            s = "".join("ACGTN"[c] for c in nt4)
            fh.write(f">{name} some description\n")
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")


def test_bam_reader_roundtrip(tmp_path, plp_case):
    _need(BISCUIT)
    ref, rd = plp_case
    small = synth.make_reference(40_000, 1, seed=5)[0][1]
    rd2 = synth_plp.make_reads(small, 300, seed=4, noise=True)
    bam = str(tmp_path / "t.bam")
    # small blocks so that records straddle BGZF blocks and the linear index has several entries
    n = bamio.write_bam_from_soa(bam, [("chrA", len(ref)), ("chrEmpty", 1000), ("chrB", len(small))], [rd, None, rd2], block=3000, tag_style="ZS")
    assert n == rd["n_reads"] + rd2["n_reads"]
    out = subprocess.run([BISCUIT, "bamdump", bam], capture_output=True, check=True).stdout.decode().splitlines()
    hdr = [l for l in out if l.startswith("@")]
    assert hdr == [f"@\tchrA\t{len(ref)}", "@\tchrEmpty\t1000", f"@\tchrB\t{len(small)}"]
    rows = [l.split("\t") for l in out if not l.startswith("@")]
    assert len(rows) == n
    k = 0
    for tid, r in ((0, rd), (2, rd2)):
        for i in range(int(r["n_reads"])):
            f = rows[k]
            k += 1
            nc, co = int(r["n_cigar"][i]), int(r["cigar_off"][i])
            exp = [tid, r["pos"][i], r["mpos"][i], r["flag"][i], r["mapq"][i], r["l_qseq"][i], r["nm"][i], r["as_"][i], r["mate_rlen"][i],
                   r["bss_tag"][i], nc] + r["cigar"][co:co + nc].tolist()
            assert [int(x) for x in f[:-1]] == [int(x) for x in exp], (tid, i)
    # BAI: seeking to chrB / a late window of chrA yields exactly the records from there on
    sub = subprocess.run([BISCUIT, "bamdump", bam, "2"], capture_output=True, check=True).stdout.decode().splitlines()
    assert len([l for l in sub if not l.startswith("@")]) == rd2["n_reads"]
    sub = subprocess.run([BISCUIT, "bamdump", bam, "0", "200000"], capture_output=True, check=True).stdout.decode().splitlines()
    got_pos = [int(l.split("\t")[1]) for l in sub if not l.startswith("@")]
    need = [int(p) for p in rd["pos"] if p + 160 > 200000]
    assert got_pos[-len(need):] == need and len(got_pos) < rd["n_reads"]
    assert subprocess.run([BISCUIT, "bamdump", bam, "1"], capture_output=True, check=True).stdout == b""


def _py_vcf2bed(vcf_text, target="CG", k=1):
    out = []
    for line in vcf_text.splitlines():
        if line.startswith("#"):
            continue
        f = line.split("\t")
        info = dict(x.split("=", 1) for x in f[7].split(";") if "=" in x)
        if "CX" not in info:
            continue
        if target == "C":
            if f[3] not in "CG":
                continue
        elif target == "CH":
            if info["CX"] not in ("CHH", "CHG"):
                continue
        elif info["CX"] != target:
            continue
        keys = f[8].split(":")
        vals = f[9].split(":")
        d = dict(zip(keys, vals))
        cov = int(d["CV"]) if d.get("CV", ".") not in (".",) else 0
        if cov < k:
            continue
        bt = d.get("BT", ".")
        out.append(f"{f[0]}\t{int(f[1]) - 1}\t{f[1]}\t" + ("." if bt == "." else "%1.3f" % float(bt)) + f"\t{cov}")
    return "\n".join(out) + ("\n" if out else "")


def test_vcf2bed_and_mergecg(tmp_path, plp_case):
    _need(oracle_plp.SO, BISCUIT)
    ref, rd = plp_case
    rd1 = dict(rd)
    rd1["sid"] = np.zeros_like(rd["sid"])
    recs = oracle_plp.region(oracle_plp.conf_default(), ref, rd1, 1, len(ref), 1)
    text, _, _ = oracle_vcf(recs, "chrT", 1)
    vcf = tmp_path / "o.vcf"
    vcf.write_text("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\n" + text.decode())
    for target, k in (("cg", 1), ("ch", 3), ("c", 1)):
        got = subprocess.run([BISCUIT, "vcf2bed", "-t", target, "-k", str(k), str(vcf)], capture_output=True, check=True).stdout.decode()
        assert got == _py_vcf2bed(vcf.read_text(), target.upper(), k)
        assert len(got) > 100
    ectx = subprocess.run([BISCUIT, "vcf2bed", "-e", str(vcf)], capture_output=True, check=True).stdout.decode().splitlines()[0].split("\t")
    assert ectx[3] in "CG" and ectx[4] == "CG" and len(ectx[6]) == 5 and ectx[5] == ectx[6][2:4]
    snp = subprocess.run([BISCUIT, "vcf2bed", "-t", "snp", str(vcf)], capture_output=True, check=True).stdout.decode()
    for line in snp.splitlines():
        f = line.split("\t")
        assert len(f) == 9 and f[4] != "." and float(f[8]) > 0
    # mergecg: C and G of one CpG become one row [C.pos-1, G.pos]; betas are re-derived from the counts
    bed = tmp_path / "cg.bed"
    bed.write_text(subprocess.run([BISCUIT, "vcf2bed", "-t", "cg", str(vcf)], capture_output=True, check=True).stdout.decode())
    fa = tmp_path / "r.fa"
    _write_fasta(str(fa), [("chrT", ref)])
    merged = subprocess.run([BISCUIT, "mergecg", str(fa), str(bed)], capture_output=True, check=True).stdout.decode().splitlines()
    rows = [l.split("\t") for l in bed.read_text().splitlines()]
    by_end = {int(r[2]): r for r in rows}
    n_pairs = 0
    for m in merged:
        f = m.split("\t")
        beg, end = int(f[1]), int(f[2])
        assert end - beg == 2 and ref[beg] == 1 and ref[beg + 1] == 2  # synthetic reference: every CG row sits in a CpG
        c, g = by_end.get(beg + 1), by_end.get(end)
        cd, gd = (int(c[4]) if c else 0), (int(g[4]) if g else 0)
        M = (round(float(c[3]) * cd) if c else 0) + (round(float(g[3]) * gd) if g else 0)
        assert f[3] == "%1.3f" % (M / (cd + gd)) and int(f[4]) == cd + gd
        assert f[5] == ("C:.:0" if not c else "C:%1.3f:%d" % (float(c[3]), cd)) + "," + ("G:.:0" if not g else "G:%1.3f:%d" % (float(g[3]), gd))
        n_pairs += bool(c and g)
    assert n_pairs > 50 and len(merged) == len(rows) - n_pairs


def _oracle_cli_expected(contigs, reads_by_contig, n_bams, step=100000, chunk_beg=None):
    """VCF body + meth_average rows via the oracle, contigs in name order."""
    conf = oracle_plp.conf_default()
    body, stats = [], {}
    for name, nt4 in sorted(contigs, key=lambda c: c[0]):
        rd = reads_by_contig.get(name)
        if rd is None:
            continue
        recs = oracle_plp.region(conf, nt4, rd, 1, len(nt4), n_bams)
        text, beta, cnt = oracle_vcf(recs, name, n_bams, step=step)
        body.append(text)
        stats[name] = (beta, cnt)
    return b"".join(body), stats


@pytest.mark.gpu
@pytest.mark.parametrize("n_bams", [1, 2])
def test_pileup_cli_matches_oracle(tmp_path, n_bams):
    _need(oracle_plp.SO, BISCUIT)
    _check_pileup_cli(BISCUIT, tmp_path, n_bams)


@pytest.mark.parametrize("n_bams", [1, 2])
def test_pileup_cli_host_side(tmp_path, n_bams):
    """The same command-line check without a GPU: the host program linked against the test-only emulation, whose
    pileup half is the oracle's restatement.  Covers the host code around the kernels (BGZF / BAM / BAI / FASTA
    readers, chunking and carry-over, VCF text, methylation averages), not the kernels."""
    _need(oracle_plp.SO)
    import test_align_sam
    _check_pileup_cli(test_align_sam.build_emu_bin(), tmp_path, n_bams)


def _check_pileup_cli(BISCUIT, tmp_path, n_bams):
    ref_b = synth.make_reference(250_000, 1, seed=3, n_runs=2)[0][1]
    ref_a = synth.make_reference(120_000, 1, seed=8)[0][1]
    ref_c = synth.make_reference(30_000, 1, seed=9)[0][1]
    # BAM header order chr2, chr10, chr1: output must come in name order chr1, chr10, chr2; chr1 has no reads
    contigs = [("chr2", ref_b), ("chr10", ref_a), ("chr1", ref_c)]
    rd_b = synth_plp.make_reads(ref_b, 5000, seed=9, noise=True, n_bams=n_bams)
    rd_a = synth_plp.make_reads(ref_a, 2500, seed=10, noise=True, n_bams=n_bams)
    fa = str(tmp_path / "ref.fa")
    _write_fasta(fa, contigs)
    bams = []
    for s in range(n_bams):
        b = str(tmp_path / f"s{s}.bam")
        bamio.write_bam_from_soa(b, [(n, len(x)) for n, x in contigs], [rd_b, rd_a, None], sid=s, block=20000,
                                 tag_style=("YD", "XG")[s % 2])
        bams.append(b)
    out = str(tmp_path / "out.vcf")
    # a step that does not divide the contig and several chunks per contig (chunk = whole windows <= 8 M loci)
    subprocess.run([BISCUIT, "pileup", "-@", "4", "-s", "70000", "-o", out, fa] + bams, check=True)
    got = open(out, "rb").read()
    hdr = [l for l in got.split(b"\n") if l.startswith(b"#")]
    assert hdr[0] == b"##fileformat=VCFv4.1" and hdr[-1].split(b"\t")[9:] == [f"s{s}".encode() for s in range(n_bams)]
    assert [l for l in hdr if l.startswith(b"##contig")] == [b"##contig=<ID=chr1,length=30000>", b"##contig=<ID=chr10,length=120000>",
                                                              b"##contig=<ID=chr2,length=250000>"]
    body = b"\n".join(l for l in got.split(b"\n") if not l.startswith(b"#"))
    exp, stats = _oracle_cli_expected(contigs, {"chr2": rd_b, "chr10": rd_a}, n_bams, step=70000)
    assert body == exp
    # <out>_meth_average.tsv: CG / CHG / CHH / CH counts and means per contig + genome, %1.3f%% of double sums
    tsv = open(out + "_meth_average.tsv").read().splitlines()
    assert tsv[0] == "sample\tchrm\tCGn\tCGb\tCHGn\tCHGb\tCHHn\tCHHb\tCHn\tCHb"

    def row(sample, label, b, c):
        k_cg, k_chg, k_chh = c[3] + c[0], c[4] + c[1], c[5] + c[2]
        b_cg, b_chg, b_chh = b[3] + b[0], b[4] + b[1], b[5] + b[2]
        k_ch, b_ch = k_chg + k_chh, b_chg + b_chh
        return (f"{sample}\t{label}\t{k_cg}\t%1.3f%%\t{k_chg}\t%1.3f%%\t{k_chh}\t%1.3f%%\t{k_ch}\t%1.3f%%"
                % (b_cg / k_cg * 100, b_chg / k_chg * 100, b_chh / k_chh * 100, b_ch / k_ch * 100))

    # row k = sums of BAM contig k, labelled sorted[sorted[k].tid].name (src/pileup.c:121-143):
    # header order (chr2, chr10, chr1) -> name order (chr1:tid2, chr10:tid1, chr2:tid0)
    sorted_t = [("chr1", 2), ("chr10", 1), ("chr2", 0)]
    exp_rows = []
    for s in range(n_bams):
        tot_b, tot_c = np.zeros(6), np.zeros(6, np.int64)
        for k, name_of_tid_k in enumerate(["chr2", "chr10", "chr1"]):
            if name_of_tid_k not in stats:
                continue
            b, c = stats[name_of_tid_k][0][s * 6:s * 6 + 6], stats[name_of_tid_k][1][s * 6:s * 6 + 6]
            label = sorted_t[sorted_t[k][1]][0]
            exp_rows.append(row(bams[s], label, b, c))
            tot_b += b
            tot_c += c
        exp_rows.append(row(bams[s], "WholeGenome", tot_b, tot_c))
    assert tsv[1:] == exp_rows
    # region mode
    reg = subprocess.run([BISCUIT, "pileup", "-g", "chr2:100,001-150000", fa] + bams, capture_output=True, check=True).stdout
    rbody = b"\n".join(l for l in reg.split(b"\n") if not l.startswith(b"#"))
    recs = oracle_plp.region(oracle_plp.conf_default(), ref_b, rd_b, 100001, 150000, n_bams)
    assert rbody == oracle_vcf(recs, "chr2", n_bams, w0=100001)[0]


@pytest.mark.gpu
def test_align_to_pileup_end_to_end(tmp_path):
    """BASELINE.json configs[4] in miniature: `biscuit index` -> `biscuit align` (GPU) -> `biscuit sortbam` ->
    `biscuit pileup` (GPU) -> `vcf2bed`.  The SAM must equal the reference's, and the VCF must equal what the oracle
    derives from that SAM (tags YD / NM / AS / MC as the aligner wrote them)."""
    _need(oracle_plp.SO, BISCUIT)
    _check_end_to_end(BISCUIT, BISCUIT, tmp_path, 12000)  # about 24x


def test_align_to_pileup_end_to_end_host_side(tmp_path):
    """The same chain on a machine without a GPU: the host programs linked against the test-only emulation (the index
    comes from the reference's `index`, which the emulation cannot build).  Checks the host code of every step --
    phase 2 of the aligner, sortbam, the BAM readers and VCF text of pileup, vcf2bed -- not the kernels."""
    import refprobe
    import test_align_sam
    _need(oracle_plp.SO)
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    _check_end_to_end(test_align_sam.build_emu_bin(), refprobe.REF_BIN, tmp_path, 2500)  # about 5x


def _check_end_to_end(BISCUIT, INDEXER, tmp_path, n_pairs):
    import refprobe
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    ref = synth.make_reference(150_000, 2, seed=21)
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta(fa, ref)
    subprocess.run([INDEXER, "index", fa], check=True, capture_output=True)
    p = synth.simulate_pairs(ref, n_pairs, seed=31, sub_rate=0.01, indel_rate=0.001, qual="mixed")
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, p["r1"], p["q1"], suffix="/1")
    synth.write_fastq(f2, p["r2"], p["q2"], suffix="/2")
    sam = str(tmp_path / "out.sam")
    with open(sam, "wb") as fh:
        subprocess.run([BISCUIT, "align", "-@", "4", fa, f1, f2], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    ref_sam = subprocess.run([refprobe.REF_BIN, "align", "-@", "4", fa, f1, f2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    strip = lambda b: b"\n".join(ln for ln in b.split(b"\n") if not ln.startswith(b"@PG"))  # noqa: E731
    assert strip(open(sam, "rb").read()) == strip(ref_sam)
    bam = str(tmp_path / "out.bam")
    subprocess.run([BISCUIT, "sortbam", "-@", "4", "-o", bam, sam], check=True, capture_output=True)
    vcf = str(tmp_path / "out.vcf")
    subprocess.run([BISCUIT, "pileup", "-@", "4", "-o", vcf, fa, bam], check=True)
    body = b"\n".join(ln for ln in open(vcf, "rb").read().split(b"\n") if not ln.startswith(b"#"))
    contigs, soa = bamio.sam_to_soa(sam)
    nt4 = dict(ref)
    exp = []
    for name, _ in sorted(contigs):
        if name not in soa:
            continue
        recs = oracle_plp.region(oracle_plp.conf_default(), nt4[name], soa[name], 1, len(nt4[name]), 1)
        exp.append(oracle_vcf(recs, name, 1)[0])
    assert body == b"".join(exp)
    assert body.count(b"\n") > (20000 if n_pairs >= 12000 else 5000)  # most cytosines of 150 kb are covered at 24x
    bed = subprocess.run([BISCUIT, "vcf2bed", "-t", "cg", vcf], check=True, capture_output=True).stdout.decode()
    assert bed == _py_vcf2bed(open(vcf).read(), "CG", 1) and bed.count("\n") > (1000 if n_pairs >= 12000 else 300)


def test_sortbam_matches_python_writer(tmp_path):
    """`biscuit sortbam` (C: SAM -> coordinate-sorted BAM + BAI) against tools/bamio.py on the reference aligner's own
    SAM: same decoded records in the same order, and the BAI lands a reader on the same records."""
    import refprobe
    _need(BISCUIT)
    if not refprobe.available():
        pytest.skip("oracle/_ref not built")
    ref = synth.make_reference(90_000, 3, seed=12, n_runs=1)
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta(fa, ref)
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = synth.simulate_pairs(ref, 1500, seed=2, sub_rate=0.02, indel_rate=0.004, qual="mixed")
    r2 = p["r2"].copy()
    r2[:40] = np.random.default_rng(0).integers(0, 4, size=(40, 150))  # some unmapped mates
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(f1, p["r1"], p["q1"], suffix="/1")
    synth.write_fastq(f2, r2, p["q2"], suffix="/2")
    sam = str(tmp_path / "a.sam")
    with open(sam, "wb") as fh:
        subprocess.run([refprobe.REF_BIN, "align", "-@", "2", fa, f1, f2], check=True, stdout=fh, stderr=subprocess.DEVNULL)
    py_bam, c_bam = str(tmp_path / "py.bam"), str(tmp_path / "c.bam")
    n = bamio.sam_to_sorted_bam(sam, py_bam)
    subprocess.run([BISCUIT, "sortbam", "-@", "3", "-o", c_bam, sam], check=True, capture_output=True)
    dump = lambda *a: subprocess.run([BISCUIT, "bamdump", *a], capture_output=True, check=True).stdout  # noqa: E731
    a, b = dump(py_bam), dump(c_bam)
    assert a == b and a.count(b"\n") > n * 0.9
    for args in (("1",), ("2", "20000"), ("0", "5000")):
        assert dump(py_bam, *args) == dump(c_bam, *args)
    # the file is a valid multi-member gzip stream holding a BAM
    import gzip
    raw = gzip.open(c_bam, "rb").read()
    assert raw[:4] == b"BAM\x01"
