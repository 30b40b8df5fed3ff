"""N>1 path on CPU: two ranks (torch.distributed, gloo, 127.0.0.1) shard the (read, conversion) tasks, each aligns its
shard through the C ABI (host emulation of libbsq here, libbsq.so on the GPU box), no data-path collective; the
gathered result must equal the single-rank result (reads shard embarrassingly)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lib_path, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
    import torch
    import torch.distributed as dist
    import golden_io
    import refprobe
    from biscuit_b200 import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bsq = capi.Bsq(lib_path)
    hi, z = golden_io.load_align_tiny()
    dx = bsq.upload(hi)
    al = capi.Aligner(dx, bsq.default_opt())
    n = len(z["lens"])
    seqs = np.concatenate([z["seqs"], z["seqs"]])
    lens = np.concatenate([z["lens"], z["lens"]])
    par = np.concatenate([np.zeros(n, np.uint8), np.ones(n, np.uint8)])
    lo, hi_ = (2 * n * rank) // world, (2 * n * (rank + 1)) // world  # contiguous shard of the task list
    regs, off = al.phase1(seqs[lo:hi_], lens[lo:hi_], par[lo:hi_])
    mine = refprobe.regs_from_bsq(regs)
    counts = torch.tensor([len(mine), hi_ - lo], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, counts)  # only sizes travel; the data path has no collective
    np.save(os.path.join(out_dir, f"regs{rank}.npy"), mine)
    np.save(os.path.join(out_dir, f"off{rank}.npy"), off)
    dist.barrier()
    if rank == 0:
        allr = np.concatenate([np.load(os.path.join(out_dir, f"regs{r}.npy")) for r in range(world)])
        assert int(sum(int(g[0]) for g in gathered)) == len(allr)
        assert (allr == z["regs"]).all()
        offs = [np.load(os.path.join(out_dir, f"off{r}.npy")) for r in range(world)]
        merged = [0]
        for o in offs:
            merged += (o[1:] + merged[-1]).tolist()
        assert (np.array(merged) == z["reg_off"]).all()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    import conftest
    lib = conftest.build_hostemu()
    mp.spawn(_worker, args=(2, _free_port(), lib, str(tmp_path)), nprocs=2, join=True)


def _plp_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
    import torch.distributed as dist
    import oracle_plp
    import synth
    import synth_plp
    from biscuit_b200 import plp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # two contigs, one per rank (pileup shards by contig / window; records are per-rank, only the statistics meet)
    contigs = [synth.make_reference(60_000, 1, seed=11 + i)[0][1] for i in range(world)]
    reads = [synth_plp.make_reads(c, 1200, seed=21 + i, noise=True) for i, c in enumerate(contigs)]
    conf = oracle_plp.conf_default()
    cnt = np.zeros((world, 1, 6), np.int64)
    beta = np.zeros((world, 1, 6))
    recs = oracle_plp.region(conf, contigs[rank], reads[rank], 1, len(contigs[rank]), 1)  # stands in for the GPU records (same layout)
    cnt[rank], beta[rank] = plp.context_stats(recs, 1)
    mc, mb = plp.merge_stats(cnt, beta)
    if rank == 0:
        exp_c = np.zeros_like(cnt)
        exp_b = np.zeros_like(beta)
        for i in range(world):
            r = oracle_plp.region(conf, contigs[i], reads[i], 1, len(contigs[i]), 1)
            exp_c[i], exp_b[i] = plp.context_stats(r, 1)
        assert (mc == exp_c).all() and mc.sum() > 1000
        assert mb.tobytes() == exp_b.tobytes()
        # and context_stats agrees with the reference-order accumulation of the oracle's text restatement
        import test_pileup_cli
        _, ob, oc = test_pileup_cli.oracle_vcf(oracle_plp.region(conf, contigs[0], reads[0], 1, len(contigs[0]), 1), "c", 1)
        assert oc.tolist() == exp_c[0, 0].tolist() and ob.tobytes() == exp_b[0, 0].tobytes()
    dist.barrier()
    dist.destroy_process_group()


def test_pileup_stats_reduce_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    import oracle_plp
    if not oracle_plp.available():
        pytest.skip("oracle not built")
    mp.spawn(_plp_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
