"""The drop-in boundary proven inside the reference's own source tree: the reference's
lib/aln compiled as it is, except that step 1 of mem_process_seqs (lib/aln/bwamem.c:432-476) is ONE call into the
bsq.h C ABI (oracle/ref_glue/bwamem_gpu.c #includes the reference's bwamem.c where it lies and redefines only that
function).  Everything around it -- option parsing, FASTQ reader, read clipping, mem_merge_regions, mem_pestat,
bis_worker2 with pairing / mapQ / CIGAR / SAM text, the output loop -- is the reference's own code.  The SAM must equal
the unmodified reference's.

* gpu:     oracle/_ref/biscuit_ref_gpu (links biscuit_b200/csrc/libbsq.so, CUDA)
* not gpu: the same objects linked against the test-only host emulation of the ABI (checks the glue, not the kernels)"""
import glob
import os
import subprocess

import pytest

import refprobe
from test_align_sam import _sam, hard_set, repeat_set  # noqa: F401  (fixtures)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "biscuit_ref_gpu")
REF_EMU = os.path.join(ROOT, "tests", "hostemu", "biscuit_ref_emu")

CASES = [[], ["-b", "1"], ["-5", "4", "-3", "7", "-z", "15"], ["-A", "2"], ["-I", "450,40"], ["-M", "-Y"]]


def build_ref_emu():
    import conftest
    lib = conftest.build_hostemu()
    objs = [o for o in glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "*.o"))
            if os.path.basename(o) not in ("bwamem.o", "ref_probe.o")]
    if not any(o.endswith("bwamem_gpu.o") for o in objs):
        pytest.skip("oracle/_ref/obj/bwamem_gpu.o not built (needs /root/reference at build time)")
    if not os.path.exists(REF_EMU) or any(os.path.getmtime(d) > os.path.getmtime(REF_EMU) for d in objs + [lib]):
        subprocess.check_call(["gcc", "-o", REF_EMU] + objs + ["-L" + os.path.dirname(lib), "-lbsq_hostemu", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread", "-lm"])
    return REF_EMU


@pytest.mark.parametrize("extra", CASES[:3], ids=[" ".join(c) or "default" for c in CASES[:3]])
def test_reference_with_abi_step1_hostemu(hard_set, extra):
    fa, f1, f2 = hard_set
    args = ["-@", "4"] + extra + [fa, f1, f2]
    assert _sam(build_ref_emu(), args) == _sam(refprobe.REF_BIN, args)


def test_reference_with_abi_step1_single_end_hostemu(hard_set):
    fa, f1, _ = hard_set
    args = ["-@", "2", fa, f1]
    assert _sam(build_ref_emu(), args) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", CASES, ids=[" ".join(c) or "default" for c in CASES])
def test_reference_with_abi_step1_gpu(hard_set, extra):
    if not os.path.exists(REF_GPU):
        pytest.skip("oracle/_ref/biscuit_ref_gpu not built")
    fa, f1, f2 = hard_set
    args = ["-@", "4"] + extra + [fa, f1, f2]
    assert _sam(REF_GPU, args) == _sam(refprobe.REF_BIN, args)


@pytest.mark.gpu
def test_reference_with_abi_step1_repeats_gpu(repeat_set):
    if not os.path.exists(REF_GPU):
        pytest.skip("oracle/_ref/biscuit_ref_gpu not built")
    fa, f1, f2 = repeat_set
    args = ["-@", "3", fa, f1, f2]
    assert _sam(REF_GPU, args) == _sam(refprobe.REF_BIN, args)
    args = ["-@", "3", fa, f1]
    assert _sam(REF_GPU, args) == _sam(refprobe.REF_BIN, args)
