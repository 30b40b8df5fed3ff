"""pytest plumbing: markers, synthetic data sets and the two ABI implementations under test.

* `cuda`    = biscuit_b200/csrc/libbsq.so (the product; tests marked `gpu`)
* `hostemu` = tests/hostemu (the same per-task device code compiled for the CPU; test-only)
The checker is always the oracle: oracle/_ref (the unmodified reference compiled by
oracle/Makefile) and/or the committed golden vectors under tests/golden.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def build_hostemu() -> str:
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    out = os.path.join(ROOT, "tests", "hostemu", "libbsq_hostemu.so")
    cs = os.path.join(ROOT, "biscuit_b200", "csrc")
    # the pileup half of the emulation is the oracle's restatement (test-only library: it may link the oracle)
    orc = os.path.join(ROOT, "oracle", "bsq_oracle_pileup.c")
    dpc = os.path.join(ROOT, "tests", "hostemu", "hostemu_dp.c")
    core = os.path.join(ROOT, "biscuit_b200", "host", "bq_core.c")
    deps = [src, orc, dpc, core, os.path.join(ROOT, "tests", "hostemu", "bsq_ksw_scalar.h"), os.path.join(ROOT, "biscuit_b200", "host", "bq.h"), os.path.join(ROOT, "oracle", "bsq_oracle.h"),
            os.path.join(ROOT, "include", "bsq.h")]
    deps += [os.path.join(cs, f) for f in os.listdir(cs) if f.endswith((".h", ".cuh"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        obj = os.path.join(ROOT, "tests", "hostemu", "bsq_oracle_pileup.o")
        subprocess.check_call(["gcc", "-O2", "-g", "-std=gnu11", "-fPIC", "-c", "-o", obj, orc])
        # the phase-2 DP entry points (bsq_dp_*) answered by the host's scalar routines (test-only stand-in for bsq_dp.cu)
        objs = [obj]
        for c in (dpc, core):
            o = os.path.join(ROOT, "tests", "hostemu", os.path.basename(c)[:-2] + ".o")
            subprocess.check_call(["gcc", "-O2", "-g", "-std=gnu11", "-fPIC", "-c", "-o", o, c])
            objs.append(o)
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src] + objs + ["-lm"])
    return out


@pytest.fixture(scope="session")
def hostemu():
    from biscuit_b200 import capi
    return capi.Bsq(build_hostemu())


@pytest.fixture(scope="session")
def cuda():
    from biscuit_b200 import capi
    return capi.load()  # raises if libbsq.so is missing: no fallback


def _make_dataset(tmp, name, total_len, n_contigs, n_pairs, seed, **kw):
    import refprobe
    import synth
    if not refprobe.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    d = tmp.mktemp(name)
    fa = os.path.join(str(d), "ref.fa")
    ref = synth.make_reference(total_len, n_contigs, seed=7, n_runs=kw.pop("n_runs", 0))
    synth.write_fasta(fa, ref)
    subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    pairs = synth.simulate_pairs(ref, n_pairs, seed=seed, **kw)
    return dict(fa=fa, ref=ref, pairs=pairs, dir=str(d))


@pytest.fixture(scope="session")
def ds_1m(tmp_path_factory):
    """1 Mb reference, 2 contigs, clean 2x150 reads (BASELINE.json configs[0] shape)."""
    return _make_dataset(tmp_path_factory, "ds1m", 1_000_000, 2, 600, 1)


@pytest.fixture(scope="session")
def ds_hard(tmp_path_factory):
    """Small reference (200 kb incl. N runs, 5 contigs) with noisy reads: substitutions 2 %, indels, N bases,
    mixed qualities -- exercises ties, re-seeding, band doubling and contig edges."""
    return _make_dataset(tmp_path_factory, "dshard", 200_037, 5, 400, 3, sub_rate=0.02, indel_rate=0.004, n_rate=0.002,
                         qual="mixed", n_runs=3)
