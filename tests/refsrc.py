"""Handle on oracle/_ref/biscuit_ref_src: the UNMODIFIED reference src/pileup.c, src/bisc_utils.c, src/vcf2bed.c and
src/mergecg.c compiled over the header stand-ins of oracle/ref_shim_src (oracle/Makefile).  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC_BIN = os.path.join(ROOT, "oracle", "_ref", "biscuit_ref_src")

# VCF columns that depend on huishenlab/utils stats.h (absent from the reference tree): QUAL, FILTER and the GT / GL1 / GQ
# sub-fields.  Both sides print them from the same restated formulas, so whole lines are compared; `pinned_view` is the
# projection that holds even if that restatement were wrong.
UNPINNED_FORMAT_KEYS = ("GT", "GL1", "GQ")


def available() -> bool:
    return os.path.exists(REF_SRC_BIN)


def run(*args, **kw):
    return subprocess.run([REF_SRC_BIN, *args], capture_output=True, check=True, **kw)


def split_vcf(data: bytes):
    """-> (header lines without ##program / ##source, body lines)."""
    lines = data.split(b"\n")
    hdr = [l for l in lines if l.startswith(b"#") and not l.startswith(b"##program") and not l.startswith(b"##source")]
    body = [l for l in lines if l and not l.startswith(b"#")]
    return hdr, body


def pinned_view(body):
    """Every column that is a function of code present under /root/reference: drops QUAL, FILTER, GT, GL1, GQ."""
    out = []
    for l in body:
        f = l.split(b"\t")
        keys = f[8].split(b":")
        keep = [i for i, k in enumerate(keys) if k.decode() not in UNPINNED_FORMAT_KEYS]
        cols = f[:5] + [f[7], b":".join(keys[i] for i in keep)]
        for smp in f[9:]:
            v = smp.split(b":")
            cols.append(b":".join(v[i] for i in keep if i < len(v)))
        out.append(b"\t".join(cols))
    return out
