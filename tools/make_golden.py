"""Generate tests/golden/align_tiny.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only (needs oracle/_ref):

    python tools/make_golden.py

The fixture holds a 30 kb / 3 contig synthetic reference, its index exactly as `biscuit_ref index`
wrote it, 64 simulated 2x150 bisulfite pairs (+ ragged extras) and, for every (read, conversion)
task, what the reference computes: the SMEM interval list (mem_collect_intv), the filtered chains
(mem_chain + mem_chain_flt) and the regions of mem_align1_core.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refprobe  # noqa: E402
import synth  # noqa: E402
from biscuit_b200 import indexio  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden", "align_tiny.npz")
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "tiny.fa")
        ref = synth.make_reference(30_000, 3, seed=7, n_runs=1)
        synth.write_fasta(fa, ref)
        subprocess.check_call([refprobe.REF_BIN, "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        pairs = synth.simulate_pairs(ref, 64, seed=42, sub_rate=0.01, indel_rate=0.002, n_rate=0.001)
        reads = [np.asarray(r, np.uint8) for r in pairs["r1"]] + [np.asarray(r, np.uint8) for r in pairs["r2"]]
        reads += [reads[0][:40], reads[1][:18], reads[2][:19], reads[3][:1], np.full(30, 4, np.uint8)]
        hi = indexio.load_index(fa)
        rp = refprobe.RefProbe(fa)
        n = len(reads)
        L = max(len(r) for r in reads)
        mat = np.zeros((n, L), np.uint8)
        lens = np.array([len(r) for r in reads], np.int32)
        for i, r in enumerate(reads):
            mat[i, :len(r)] = r
        intv, intv_off, chains, chain_off, regs, reg_off = [], [0], [], [0], [], [0]
        for parent in (0, 1):
            for r in reads:
                iv = rp.collect_intv(parent, r) if len(r) >= 19 else np.zeros((0, 4), np.uint64)
                intv.append(iv)
                intv_off.append(intv_off[-1] + len(iv))
                ch, _, _ = rp.chain(parent, r, stage=1)
                chains.append(ch)
                chain_off.append(chain_off[-1] + len(ch))
                rg = refprobe.regs_from_ref(rp.align1(parent, r))
                regs.append(rg)
                reg_off.append(reg_off[-1] + len(rg))
        rp.close()
        np.savez_compressed(
            out, seqs=mat, lens=lens,
            intv=np.concatenate(intv), intv_off=np.array(intv_off), chains=np.concatenate(chains),
            chain_off=np.array(chain_off), regs=np.concatenate(regs), reg_off=np.array(reg_off),
            bwt0=hi.fm[0].bwt, bwt1=hi.fm[1].bwt, sa0=hi.fm[0].sa, sa1=hi.fm[1].sa,
            meta=np.array([hi.fm[0].primary, hi.fm[1].primary, hi.fm[0].sa_intv, hi.fm[0].seq_len, hi.l_pac], np.int64),
            L2=np.stack([hi.fm[0].L2, hi.fm[1].L2]), pac=hi.pac, ann_offset=hi.ann_offset, ann_len=hi.ann_len,
            names=np.array(hi.names))
    print("wrote", out, os.path.getsize(out), "bytes;", n, "reads,", reg_off[-1], "regions")


if __name__ == "__main__":
    main()
