"""Freeze a small pileup case whose expected output comes from the REFERENCE ITSELF (oracle/_ref/biscuit_ref_src =
the unmodified src/pileup.c, src/vcf2bed.c, src/mergecg.c over oracle/ref_shim_src) into tests/golden/pileup_tiny/.

    python tools/make_golden_pileup.py        # needs oracle/_ref (i.e. /root/reference at build time)

Inputs: ref.fa.gz (3 contigs, header order != name order), s0.bam/s1.bam (+ .bai) with YD resp. XG strand tags, reads
with indel / soft-clip / hard-clip CIGARs, filter-triggering flags and missing tags.  Outputs: VCF + _meth_average.tsv for
two option sets, the CpG BED of vcf2bed and its mergecg.  tests/test_pileup_ref.py replays them."""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import bamio  # noqa: E402
import synth  # noqa: E402
import synth_plp  # noqa: E402
from test_pileup_cli import _write_fasta  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "biscuit_ref_src")
OUT = os.path.join(ROOT, "tests", "golden", "pileup_tiny")


def gz(src, dst):
    with open(src, "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
        fo.write(fi.read())


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp()
    ref_b = synth.make_reference(9_000, 1, seed=3, n_runs=1)[0][1]
    ref_a = synth.make_reference(5_000, 1, seed=8)[0][1]
    ref_c = synth.make_reference(2_000, 1, seed=9)[0][1]
    contigs = [("chr2", ref_b), ("chr10", ref_a), ("chr1", ref_c)]
    rd_b = synth_plp.make_reads(ref_b, 260, seed=9, noise=True, n_bams=2)
    rd_a = synth_plp.make_reads(ref_a, 140, seed=10, noise=True, n_bams=2)
    fa = os.path.join(tmp, "ref.fa")
    _write_fasta(fa, contigs)
    gz(fa, os.path.join(OUT, "ref.fa.gz"))
    bams = []
    for s in range(2):
        b = os.path.join(OUT, f"s{s}.bam")
        bamio.write_bam_from_soa(b, [(n, len(x)) for n, x in contigs], [rd_b, rd_a, None], sid=s, block=6000, tag_style=("YD", "XG")[s])
        bams.append(b)
    for tag, opts in (("default", []), ("nome_step", ["-N", "-s", "3000"])):
        v = os.path.join(tmp, tag + ".vcf")
        subprocess.run([REF, "pileup", "-o", v] + opts + [fa] + bams, check=True, capture_output=True)
        gz(v, os.path.join(OUT, tag + ".vcf.gz"))
        gz(v + "_meth_average.tsv", os.path.join(OUT, tag + ".tsv.gz"))
    v = os.path.join(tmp, "default.vcf")
    bed = subprocess.run([REF, "vcf2bed", "-t", "cg", "-s", "ALL", v], check=True, capture_output=True).stdout
    open(os.path.join(tmp, "cg.bed"), "wb").write(bed)
    gz(os.path.join(tmp, "cg.bed"), os.path.join(OUT, "cg.bed.gz"))
    mg = subprocess.run([REF, "mergecg", fa, os.path.join(tmp, "cg.bed")], check=True, capture_output=True).stdout
    open(os.path.join(tmp, "m.bed"), "wb").write(mg)
    gz(os.path.join(tmp, "m.bed"), os.path.join(OUT, "cg_merged.bed.gz"))
    shutil.rmtree(tmp)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
