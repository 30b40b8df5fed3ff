"""Pileup leg of bench.py: loci/s of the methylation caller over one chr1-sized contig per GPU at 30x.

Used two ways: `python bench.py` runs it after the align leg and carries the result as the `pileup` object of the one
JSON line; `python bench.py --path pileup` prints it as its own line.  Everything that touches oracle/ here is the
checker / the CPU baseline (oracle/_ref/biscuit_ref_src = the unmodified reference pileup), never the measured path.

  value     kernels only (bsq_plp_run over device-resident decoded reads), CUDA-event bracketed, max over ranks
  e2e       the C ABI with HOST buffers: bsq_plp_stage (H2D of every read column) + run + bsq_plp_fetch (D2H of one
            record per emitted locus) per pass
  e2e_cli   `biscuit pileup` on a BAM + FASTA of a bounded sample (N=1): BGZF inflate, BAM decode, GPU, VCF text
  cpu_baseline  the reference's own pileup (all host threads) on the same sample, and `parity`: VCF bodies compared
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
READ_LEN = 150


def make_reads_gpu(torch, nt4_dev, n_pairs: int, seed: int, chunk: int = 2_000_000):
    """Coordinate-sorted synthetic WGBS alignments generated on the device (49.6 M reads for 248 Mb at 30x take seconds;
    the numpy generator of the tests takes minutes at this size).  Fragments ~N(300,30) placed uniformly, half of them
    from the complementary bisulfite strand; read pairs FR; C->T (BSW) or G->A (BSC) conversion with 80 % retention in
    CpG and 1 % elsewhere; 0.5 % substitutions; qualities 40 with a 10 % low tail.  150M CIGARs, tags NM AS MC YD.
    Returns the structure-of-arrays the pileup ABI takes (host numpy arrays)."""
    dev = nt4_dev.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = int(nt4_dev.numel())
    flen = torch.clamp((300 + 30 * torch.randn(n_pairs, generator=g, device=dev)).round().long(), READ_LEN + 10, 600)
    pos = (torch.rand(n_pairs, generator=g, device=dev, dtype=torch.float64) * (L - 600 - 2)).long() + 1
    bsc = torch.rand(n_pairs, generator=g, device=dev) < 0.5
    left, right = pos, pos + flen - READ_LEN
    rpos = torch.cat([left, right])
    mpos = torch.cat([right, left])
    flag = torch.cat([torch.where(bsc, 163, 99), torch.where(bsc, 83, 147)]).to(torch.int32)
    bss = torch.cat([bsc, bsc])
    rpos, order = torch.sort(rpos, stable=True)
    mpos, flag, bss = mpos[order], flag[order], bss[order]
    n = int(rpos.numel())
    seq = np.empty((n, (READ_LEN + 1) // 2), np.uint8)
    qual = np.empty((n, READ_LEN), np.uint8)
    ar = torch.arange(READ_LEN, device=dev)
    nt16 = torch.tensor([1, 2, 4, 8], device=dev, dtype=torch.uint8)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        idx = rpos[c0:c1, None] + ar[None, :]
        base = nt4_dev[idx].long()
        nxt, prv = nt4_dev[idx + 1].long(), nt4_dev[idx - 1].long()
        u = torch.rand(base.shape, generator=g, device=dev)
        b = bss[c0:c1, None]
        cpg_w = (base == 1) & (nxt == 2)
        cpg_c = (base == 2) & (prv == 1)
        keep_w = u < torch.where(cpg_w, 0.8, 0.01)
        keep_c = u < torch.where(cpg_c, 0.8, 0.01)
        conv = torch.where(~b & (base == 1) & ~keep_w, 3, base)
        conv = torch.where(b & (base == 2) & ~keep_c, 0, conv)
        e = torch.rand(base.shape, generator=g, device=dev) < 0.005
        conv = torch.where(e, torch.randint(0, 4, base.shape, generator=g, device=dev), conv)
        code = nt16[conv]
        seq[c0:c1] = ((code[:, 0::2] << 4) | code[:, 1::2]).cpu().numpy()
        q = torch.full(base.shape, 40, device=dev, dtype=torch.uint8)
        low = torch.rand(base.shape, generator=g, device=dev) < 0.10
        q = torch.where(low, torch.randint(2, 40, base.shape, generator=g, device=dev).to(torch.uint8), q)
        qual[c0:c1] = q.cpu().numpy()
    ar_n = np.arange(n, dtype=np.int64)
    return dict(n_reads=n, pos=rpos.to(torch.int32).cpu().numpy(), mpos=mpos.to(torch.int32).cpu().numpy(),
                mate_rlen=np.full(n, READ_LEN, np.int32), l_qseq=np.full(n, READ_LEN, np.int32), nm=np.full(n, 1, np.int32),
                as_=np.full(n, 140, np.int32), flag=flag.cpu().numpy().astype(np.uint16), mapq=np.full(n, 60, np.uint8),
                bss_tag=bss.to(torch.int8).cpu().numpy(), sid=np.zeros(n, np.uint8), n_cigar=np.ones(n, np.int32), cigar_off=ar_n,
                cigar=np.full(n, READ_LEN << 4, np.uint32), seq=seq.reshape(-1), seq_off=ar_n * ((READ_LEN + 1) // 2), qual=qual.reshape(-1),
                qual_off=ar_n * READ_LEN)


def subset(rd, keep):
    n = rd["n_reads"]
    rs = {}
    for k, v in rd.items():
        if k == "n_reads":
            continue
        if k == "seq":
            rs[k] = v.reshape(n, -1)[keep].reshape(-1)
        elif k == "qual":
            rs[k] = v.reshape(n, -1)[keep].reshape(-1)
        else:
            rs[k] = v[keep]
    m = int(keep.sum()) if keep.dtype == bool else len(keep)
    rs["n_reads"] = m
    ar = np.arange(m, dtype=np.int64)
    rs["cigar_off"], rs["seq_off"], rs["qual_off"] = ar, ar * ((READ_LEN + 1) // 2), ar * READ_LEN
    return rs


def _vcf_body(path):
    with open(path, "rb") as fh:
        return [l for l in fh.read().split(b"\n") if l and not l.startswith(b"#")]


def sample_legs(log, nt4, rd, sample_mb: float, ncores: int, want_cpu: bool):
    """BAM + FASTA of the first `sample_mb` megabases -> (`biscuit pileup` timing, reference pileup timing, parity)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bamio
    sub = int(sample_mb * 1_000_000)
    keep = rd["pos"] < sub - 700
    rs = subset(rd, keep)
    exe = os.path.join(ROOT, "biscuit_b200", "host", "biscuit")
    refbin = os.path.join(ROOT, "oracle", "_ref", "biscuit_ref_src")
    cli = cpu = parity = None
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        fa, bam = os.path.join(d, "ref.fa"), os.path.join(d, "in.bam")
        with open(fa, "w") as fh:
            fh.write(">chrS\n")
            txt = np.frombuffer(b"ACGT", np.uint8)[nt4[:sub]].tobytes().decode()
            fh.write("\n".join(txt[i:i + 100] for i in range(0, sub, 100)) + "\n")
        bamio.write_bam_fixed(bam, "chrS", sub, rs)
        bam_bytes = os.path.getsize(bam)
        log(f"pileup sample: {rs['n_reads']} reads over {sample_mb:g} Mb, FASTA + BAM ({bam_bytes / 1e6:.0f} MB) written in {time.perf_counter() - t0:.1f}s")
        t0 = time.perf_counter()
        subprocess.run([exe, "pileup", "-@", str(ncores), "-o", os.path.join(d, "out.vcf"), fa, bam], check=True)
        dt_cli = time.perf_counter() - t0
        cli = {"value": (sub - 1) / dt_cli, "unit": "loci/s", "seconds": dt_cli, "bam_bytes": bam_bytes,
               "vcf_bytes": os.path.getsize(os.path.join(d, "out.vcf")), "host_threads": ncores,
               "sample": f"first {sample_mb:g} Mb of the contig ({rs['n_reads']} reads)",
               "note": "biscuit pileup: process start + CUDA context, FASTA load, BGZF inflate + BAM decode, GPU, VCF text, file write"}
        if want_cpu and os.path.exists(refbin):
            t0 = time.perf_counter()
            subprocess.run([refbin, "pileup", "-@", str(ncores), "-o", os.path.join(d, "ref.vcf"), fa, bam], check=True, stderr=subprocess.DEVNULL)
            dt_ref = time.perf_counter() - t0
            cpu = {"value": (sub - 1) / dt_ref, "unit": "loci/s", "cores": ncores, "kind": "reference",
                   "sample": f"first {sample_mb:g} Mb of the contig ({rs['n_reads']} reads) through oracle/_ref/biscuit_ref_src pileup "
                             "(unmodified src/pileup.c; BAM held in memory by the htslib stand-in)"}
            a, b = _vcf_body(os.path.join(d, "out.vcf")), _vcf_body(os.path.join(d, "ref.vcf"))
            parity = {"identical": a == b, "vcf_lines": len(b), "what": "VCF body of `biscuit pileup` (GPU) vs the reference's pileup on the sample BAM",
                      "tsv_identical": open(os.path.join(d, "out.vcf_meth_average.tsv"), "rb").read() ==
                      open(os.path.join(d, "ref.vcf_meth_average.tsv"), "rb").read()}
    return cli, cpu, parity


def run(args, log, torch, dist, rank: int, local_rank: int, world: int, peak: float, peak_src: str, clock_sampler_cls) -> dict | None:
    """Returns the pileup record on rank 0, None elsewhere."""
    sys.path.insert(0, ROOT)
    from biscuit_b200 import capi, plp
    dev = torch.device("cuda", local_rank)
    # Host memory: a rank keeps the read columns (57 B per locus at 30x), their page-locked copy and the page-locked output
    # (48 B per locus), about 0.18 GB per Mb of contig.  All ranks of a node share its RAM: when the requested contig
    # size does not fit in 60 % of what is available, every rank shrinks it by the same factor (rank 0 decides) and the
    # record says so -- a rank that dies of memory pressure would leave the others waiting in a collective.
    plp_mb, shrunk = float(args.plp_mb), None
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
        t = torch.tensor([avail_gb], device=dev, dtype=torch.float64)
        if world > 1:
            dist.broadcast(t, src=0)
        fit_mb = 0.6 * float(t[0]) / world / 0.18
        if fit_mb < plp_mb:
            shrunk = {"requested_mb": plp_mb, "host_ram_available_gb": float(t[0])}
            plp_mb = max(8.0, float(int(fit_mb)))
            log(f"pileup rank {rank}: contig shrunk to {plp_mb:.0f} Mb per GPU to fit the node's host memory ({float(t[0]):.0f} GB available, {world} ranks)")
    except Exception as e:  # noqa: BLE001
        log("host memory check skipped:", e)
    L = int(plp_mb * 1_000_000)
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    t0 = time.time()
    nt4_dev = torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8)
    n_pairs = int(L * args.plp_depth / (2 * READ_LEN))
    rd = make_reads_gpu(torch, nt4_dev, n_pairs, 31 + 1000 * rank)
    nt4 = nt4_dev.cpu().numpy()
    del nt4_dev
    torch.cuda.empty_cache()
    log(f"pileup rank {rank}: {rd['n_reads']} reads over {L / 1e6:.0f} Mb ({args.plp_depth}x) generated in {time.time() - t0:.1f}s")
    bsq = capi.load()
    pl = plp.Pileup(bsq, 1, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    conf = pl.default_conf()
    pl.set_contig(nt4)
    pl.stage(rd)
    n_loci = pl.run(conf, 1, L)
    for _ in range(args.warmup):
        pl.run(conf, 1, L)
    steps = max(args.steps, 12)
    sampler = clock_sampler_cls(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    kus = np.zeros(2)
    for _ in range(steps):
        pl.run(conf, 1, L)
        kus += pl.counters()[4:6]
    barrier()
    dt = time.perf_counter() - t0
    c = pl.counters()
    # --- the C ABI with host buffers: H2D of every read column + kernels + D2H of the records, per pass.  The host
    # buffers are page-locked (what `biscuit pileup` uses for its decoded batches, bq_bam.c), allocated outside the timing ---
    t_pin = time.perf_counter()
    pin_ok = 1
    try:
        rd_pin = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in rd.items()}
        out_pin = torch.empty((int(n_loci) + 1) * 88, dtype=torch.uint8).pin_memory().numpy().view(plp.REC_DTYPE)
    except Exception as e:  # noqa: BLE001
        log(f"pileup rank {rank}: page-locking the staging buffers failed ({e}); pageable buffers instead")
        pin_ok = 0
    if world > 1:  # all ranks take the same path
        tp = torch.tensor([pin_ok], device=dev, dtype=torch.int32)
        dist.all_reduce(tp, op=dist.ReduceOp.MIN)
        pin_ok = int(tp[0])
    h2d = sum(int(np.asarray(v).nbytes) for k, v in rd.items() if k != "n_reads")
    if not pin_ok:
        rd_pin = rd
        out_pin = np.zeros(int(n_loci) + 1, dtype=plp.REC_DTYPE)
    elif world > 1:
        rd = {"n_reads": rd["n_reads"]}  # the pageable copy is only needed by the command-line legs (N = 1)
    log(f"pileup rank {rank}: page-locked staging buffers ({sum(v.nbytes for v in rd_pin.values() if isinstance(v, np.ndarray)) / 1e9:.1f} + "
        f"{out_pin.nbytes / 1e9:.1f} GB) in {time.perf_counter() - t_pin:.1f}s")
    pl.stage(rd_pin)
    recs = pl.fetch_into(out_pin, pl.run(conf, 1, L))  # warm-up pass
    e2e_steps = 3
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        pl.stage(rd_pin)
        recs = pl.fetch_into(out_pin, pl.run(conf, 1, L))
    barrier()
    dt_e2e = time.perf_counter() - t1
    clocks = sampler.stop()
    kus /= steps
    reduce_ms = None
    if world > 1:  # the one collective of the path: per-contig methylation statistics -> every rank (NCCL over NVLink)
        cnt_all = np.zeros((world, 1, 6), np.int64)
        beta_all = np.zeros((world, 1, 6))
        cnt_all[rank], beta_all[rank] = plp.context_stats(recs[recs["pos"] <= 2_000_000], 1)
        barrier()
        t6 = time.perf_counter()
        cnt_m, beta_m = plp.merge_stats(cnt_all, beta_all, device=f"cuda:{local_rank}")
        barrier()
        reduce_ms = 1000 * (time.perf_counter() - t6)
        assert (cnt_m[rank] == cnt_all[rank]).all() and int((cnt_m.sum(axis=(1, 2)) > 0).sum()) == world
        tt = torch.tensor([dt, dt_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])
    del recs, out_pin, rd_pin
    pl.close()
    if rank != 0:
        return None
    ncores = os.cpu_count() or 1
    cli = cpu = parity = None
    if world == 1 and not args.no_cli:
        try:
            cli, cpu, parity = sample_legs(log, nt4, rd, args.plp_sample_mb, ncores, not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001
            log("pileup sample legs failed:", e)
    # reads (packed SEQ, QUAL, 48 B of record fields), reference base + flag per locus, one 88-byte record per emitted
    # locus; the per-locus counters stay in shared memory
    alg = rd["n_reads"] * (75 + 150 + 48) + (L - 1) * (1 + 4) + n_loci * 88
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary_r02.json")) as fh:
            kk = json.load(fh)["kernels"].get("k_plp_win")
        if kk and "dram_bytes_per_locus" in kk:
            traffic = kk["dram_bytes_per_locus"] * (L - 1)
    except Exception:  # noqa: BLE001
        traffic = None
    n_tiles = (L + (8 << 20) - 1) // (8 << 20)
    return {"metric": "wgbs_pileup_loci_per_s", "value": world * (L - 1) * steps / dt, "unit": "loci/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / steps, "timed_s": dt, "higher_is_better": True, "scaling": "weak", "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": f"pileup {args.plp_depth}x synthetic WGBS, coordinate-sorted decoded BAM records, one {L / 1e6:.0f} Mb "
                                   "(chr1-sized) contig per GPU, CpG/CHG/CHH extraction", "reads_per_gpu": int(rd["n_reads"]),
                       "emitted_loci_per_gpu": int(n_loci), "l2": "inputs larger than L2 (14 GB of read columns per pass)",
                       "parallelism": f"contigs sharded over {world} rank(s); one NCCL reduce of the per-contig statistics"},
            "clocks": clocks, "stats_reduce_ms": reduce_ms,
            "e2e": {"value": world * (L - 1) * e2e_steps / dt_e2e, "unit": "loci/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n_loci) * 88,
                    "steps": e2e_steps, "note": ("C ABI with page-locked host buffers" if pin_ok else "C ABI with PAGEABLE host buffers (page-locking failed)") +
                    ": bsq_plp_stage (H2D) + bsq_plp_run + bsq_plp_fetch (D2H) per pass"},
            "contig_shrunk_to_fit_host_memory": shrunk,
            "e2e_cli": cli, "parity": parity,
            "gpu_launches": 3 * steps * n_tiles,
            "roofline": {"bound": "hbm", "kernel": "k_plp_win", "achieved": alg / (kus[0] * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (kus[0] * 1e-6) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": kus[0] / 1000, "locus_kernels_ms": kus[1] / 1000, "events": int(c[3])},
            "cpu_baseline": cpu}
