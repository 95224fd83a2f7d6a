#!/usr/bin/env python
"""bench.py -- aligned reads/sec counted END TO END from a BAM file (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4] [--reads R]

Default workload = BASELINE.json configs[1]: count_reads_stranded on synthetic 100 bp reads, 25 human-like
contigs, 20k genes / ~60k exon intervals, 50M reads per GPU (weak scaling: N GPUs -> ONE N*50M-read file
sharded by chunk range).  --config selects the other named configs (0: unstranded 1M x 50 bp; 2: weighted +
introns on spliced/multi-mapper reads; 3: paired 150 bp, stranded; 4: the four-function suite).

One "step" = one complete job through the PUBLIC API, from the file on the host (page cache) to the Python
dicts: N = 1 -> the drop-in module function (mbf_bam_b200.count_reads_stranded(path, None, intervals,
gene_intervals)); N > 1 under torchrun -> mbf_bam_b200.dist (one rank per GPU, chunk-range shards, one NCCL
all-reduce of the dense vectors, every rank gets the dicts).

  value, e2e : reads/s of exactly that (host file -> H2D -> kernels -> D2H -> dicts), the same number twice.
  resident   : (extra) the same job with the compressed file already in HBM -- a kernel-pipeline figure that
               explains e2e; it is NOT the metric.
  roofline   : the kernel with the largest share of GPU time, algorithmic bytes / CUDA-event time on its stream
               against the measured HBM copy bandwidth; `roofline_all` lists every stage the same way.
  ingest     : compressed bytes over PCIe per second of e2e against the box's measured pinned H2D rate.
  cpu_baseline : the CPU restatement of the reference (oracle/: zlib + pthreads, one task per chunk like rayon in
               counters.rs:217-259) on all host cores over a bounded chunk sample of the same file.
  parity     : after the timed regions the C oracle counts the WHOLE file on all host cores and every key of
               the result dicts must be equal (merged result at N > 1).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned reads/sec counted end-to-end from BAM"
UNIT = "reads/s"


def env_int(name, default):
    v = os.environ.get(name)
    return int(v) if v else default


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch / algorithmic bytes per launch of each kernel, from the committed ncu --set full
    summary (tools/ncu_traffic.py writes profiles/ncu_traffic.json from the report); absent -> traffic null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# BASELINE.json configs -> (synth config, functions of one step, default reads per GPU, scaling)
BENCH_CONFIGS = {
    0: dict(cfg="cfg1", fns=["unstranded"], reads=1_000_000,
            what="configs[0]: count_reads_unstranded, synthetic 50bp single-end sorted+indexed BAM, 2 contigs, 1k genes"),
    1: dict(cfg="cfg2", fns=["stranded"], reads=50_000_000,
            what="configs[1]: count_reads_stranded, synthetic 100bp single-end sorted+indexed BAM, 25 human-like contigs"),
    2: dict(cfg="cfg3", fns=["weighted", "introns"], reads=100_000_000,
            what="configs[2]: count_reads_weighted + count_introns, NH-tagged multi-mappers, 35% spliced (CIGAR N) 100bp reads"),
    3: dict(cfg="cfg4", fns=["stranded"], reads=200_000_000, strong=True,
            what="configs[3]: count_reads_stranded, paired-end 150bp, per-read gene dedup, sharded by contig range"),
    4: dict(cfg="cfg5", fns=["unstranded", "stranded", "weighted", "introns"], reads=50_000_000,
            what="configs[4]: full count_reads_* suite (unstranded, stranded, weighted, introns), whole-genome-scale 100bp BAM"),
}


def dist_env():
    return env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)


def ensure_bam(bc, workdir, reads_total, rank, barrier):
    from synth import model

    c = model.CONFIGS[bc["cfg"]]
    intervals, gene_intervals, genes = model.make_gene_model(c["contigs"], c["n_genes"], c["seed"], gene_mode=c["gene_mode"],
                                                             span=c["span"], mean_exons=4.6)
    level = env_int("MBF_BENCH_LEVEL", 6)
    name = "%s_%dM_l%d.bam" % (bc["cfg"], reads_total // 1_000_000, level)

    def complete(p):
        return os.path.exists(p) and os.path.exists(p + ".bai") and os.path.exists(p + ".ok")

    # Where the synthetic file lives: the given directory, else the first of /tmp, /dev/shm, the repo with room for
    # it.  Rank 0 decides and leaves a note; the other ranks read it after the barrier.
    note = os.path.join("/tmp", "mbf_bench_%s.where" % name)
    if rank == 0:
        cands = [workdir] if workdir else ["/tmp/mbf_bench", "/dev/shm/mbf_bench", os.path.join(ROOT, ".bench")]
        need = int(reads_total * (100 if c["readlen"] > 100 else 70) * 1.1)
        chosen = next((d for d in cands if complete(os.path.join(d, name))), None)
        if chosen is None:
            for d in cands:
                try:
                    os.makedirs(d, exist_ok=True)
                    sv = os.statvfs(d)
                    if sv.f_bavail * sv.f_frsize >= need:
                        chosen = d
                        break
                except OSError:
                    continue
        if chosen is None:
            chosen = cands[0]
        with open(note, "w") as fh:
            fh.write(chosen)
        workdir = chosen
    t_gen = 0.0
    path = os.path.join(workdir or "/tmp/mbf_bench", name)
    if rank == 0 and not complete(path):
        os.makedirs(workdir, exist_ok=True)
        t0 = time.time()
        info = model.generate_bam(path, c["contigs"], genes, reads_total, c["seed"], readlen=c["readlen"], level=level,
                                  in_gene=c["in_gene"], spliced=c["spliced"], mm=c["mm"], paired=c.get("paired", False))
        t_gen = time.time() - t0
        open(path + ".ok", "w").write(json.dumps(info))
    barrier()
    if rank != 0:
        path = os.path.join(open(note).read().strip(), name)
    info = json.load(open(path + ".ok"))
    return path, intervals, gene_intervals, info, t_gen


def cpu_sample(path, intervals, gene_intervals, fns, threads, target_chunks):
    """one bounded pass of the CPU restatement over every function of the step -> (reads/s, reads, chunks, s)"""
    from oracle import c_oracle

    t0 = time.time()
    nrec = nck = 0
    for fn in fns:
        if fn == "introns":
            _, st = c_oracle.count_introns(path, nthreads=threads, return_stats=True, sample_chunks=target_chunks)
            nrec += st["n_records"]
            nck = max(nck, st["n_chunks"])
        else:
            mode = {"unstranded": 0, "stranded": 1, "weighted": 2}[fn]
            r = c_oracle.raw_counts(path, None, intervals, gene_intervals, mode, nthreads=threads, sample_chunks=target_chunks)
            nrec += r["n_records"]
            nck = max(nck, r["n_chunks"])
    dt = time.time() - t0
    return nrec / len(fns) / dt, nrec // len(fns), nck, dt


def build_checker_only():
    """the reference arm needs the BAM generator and the oracle, not the product"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def bench_config_dict(bc, args, info, intervals, reads_per_gpu):
    n_intervals = sum(len(x[2]) for v in intervals.values() for x in v)
    return {"workload": "%s, %d genes / %d intervals, zlib level %d" % (bc["what"], sum(len(v) for v in intervals.values()), n_intervals,
                                                                       env_int("MBF_BENCH_LEVEL", 6)),
            "reads": int(info["records"]), "reads_per_gpu": reads_per_gpu, "bam_bytes": int(info["bytes"]),
            "functions_per_step": bc["fns"],
            "l2": "inputs larger than L2 (compressed file %.1f GB per pass)" % (info["bytes"] / 1e9),
            "sharding": "chunk-range shards of one file, one rank per GPU" if args.gpus > 1 else "single shard"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=env_int("MBF_BENCH_CONFIG", 1), choices=sorted(BENCH_CONFIGS))
    ap.add_argument("--reads", type=int, default=env_int("MBF_BENCH_READS", 0), help="reads per GPU (config 3: in total)")
    ap.add_argument("--workdir", default=os.environ.get("MBF_BENCH_DIR"),
                    help="directory of the synthetic BAM (default: /tmp, /dev/shm or the repo, whichever has room)")
    ap.add_argument("--resident", action="store_true", help="also time the job with the file resident in HBM at N > 1")
    ap.add_argument("--no-parity", action="store_true", help="skip the whole-file oracle comparison")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    threads = os.cpu_count() or 1
    bc = BENCH_CONFIGS[args.config]
    strong = bool(bc.get("strong"))
    reads_per_gpu = args.reads or bc["reads"]
    reads_total = reads_per_gpu if strong else reads_per_gpu * max(1, args.gpus)

    if args.impl == "reference":
        # ---------------------------------------------------------------- the reference's CPU path (restated)
        # The Rust crate cannot be built here (no cargo/rustc): oracle/mbf_oracle.c, all host threads, on a
        # bounded chunk sample of the same workload per step.  Nothing of the product is imported or loaded.
        if rank != 0:
            return 0
        build_checker_only()
        path, intervals, gene_intervals, info, t_gen = ensure_bam(bc, args.workdir, reads_total, 0, lambda: None)
        from oracle import c_oracle

        n_chunks = len(c_oracle.chunks(path, None, gene_intervals))
        target = max(16, n_chunks // 6)
        vals = []
        for i in range(args.warmup + args.steps):
            v, nrec, nck, dt = cpu_sample(path, intervals, gene_intervals, bc["fns"], threads, target)
            if i >= args.warmup:
                vals.append((v, dt))
        value = statistics.mean(v for v, _ in vals)
        sample = "%d of %d chunks (%d reads) per function per step, %d threads" % (nck, n_chunks, nrec, threads)
        config = bench_config_dict(bc, args, info, intervals, reads_per_gpu)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000 * statistics.mean(dt for _, dt in vals), "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # -------------------------------------------------------------------- our arm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    else:
        # one process: the drop-in functions fan out over MBF_BAM_DEVICES (default: every visible GPU)
        os.environ.setdefault("MBF_BAM_DEVICES", ",".join(str(i) for i in range(max(1, args.gpus))))

    def barrier():
        if dist is not None:
            dist.barrier()

    import __graft_entry__ as g

    if rank == 0:
        g.build()
    barrier()
    path, intervals, gene_intervals, info, t_gen = ensure_bam(bc, args.workdir, reads_total, rank, barrier)
    n_records = int(info["records"])
    config = bench_config_dict(bc, args, info, intervals, reads_per_gpu)

    import mbf_bam_b200 as mb
    from mbf_bam_b200 import dist as mbd

    if mb.device_count() < 1:
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    fns = bc["fns"]

    def call(fn):
        """the call a user makes"""
        if dist is None:
            f = {"unstranded": mb.count_reads_unstranded, "stranded": mb.count_reads_stranded, "weighted": mb.count_reads_weighted}
            res = mb.count_introns(path) if fn == "introns" else f[fn](path, None, intervals, gene_intervals)
            return res, mb.last_stats()
        f = {"unstranded": mbd.count_reads_unstranded, "stranded": mbd.count_reads_stranded, "weighted": mbd.count_reads_weighted}
        res = mbd.count_introns(path) if fn == "introns" else f[fn](path, None, intervals, gene_intervals)
        return res, mbd.last_stats()

    def step():
        out = {}
        sts = []
        for fn in fns:
            out[fn], st = call(fn)
            sts.append(st)
        return out, sts

    def timed(stepfn, steps, warmup):
        for _ in range(warmup):
            stepfn()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        sts = []
        last = None
        for _ in range(steps):
            last, st = stepfn()
            sts.append(st)
        barrier()
        total = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None
        if dist is not None:
            import torch

            t = torch.tensor([total], dtype=torch.float64).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, sts, last, clocks

    # ---- the metric: host file -> dicts.  The first calls of the process are timed separately (cold path: copy
    #      through pinned staging buffers; after MBF_BAM_PIN_AFTER jobs the file's mapping is page-locked, see
    #      hostmap.hpp), then W warm-up steps, then K timed steps.
    t_first = []
    if dist is None:
        mb.host_cache_clear()
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        step()
        barrier()
        t_first.append(time.perf_counter() - t0)
    # (at least two more untimed steps whatever --warmup says: the page-locking of the mapping, ~0.4 s per GB once per
    #  process and file, belongs to the fourth job and must not fall into the timed region)
    e2e_total, e2e_sts, e2e_last, clocks = timed(step, args.steps, max(args.warmup, 2))

    def greduce(x, op=None):
        if dist is None:
            return x
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64).cuda()
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return float(t.item())

    def gmax(x):
        return greduce(x, dist.ReduceOp.MAX) if dist is not None else x

    flat = [s for step_sts in e2e_sts for s in step_sts]  # one entry per function call
    per_step = lambda k: sum(s[k] for s in flat) / max(1, len(e2e_sts))  # noqa: E731
    launches = int(greduce(sum(s["n_kernel_launches"] for s in flat)))
    h2d = greduce(per_step("h2d_bytes"))
    h2d_pinned = greduce(per_step("h2d_pinned_bytes"))
    d2h = greduce(per_step("d2h_bytes"))
    imbalance = None
    if dist is not None:
        mine_c, mine_ms = per_step("compressed_bytes"), per_step("ms_gpu_all")
        imbalance = {"max_over_mean_compressed_bytes": gmax(mine_c) / max(1.0, greduce(mine_c) / world),
                     "max_over_mean_gpu_ms": gmax(mine_ms) / max(1e-9, greduce(mine_ms) / world)}

    # ---- extra: compressed bytes resident in HBM (kernel pipeline only; N = 1 unless --resident)
    resident = None
    res_sts = None
    if (dist is None and args.gpus == 1) or args.resident:
        rd = mb.Reader(path)
        rd.set_device(local)
        rd.set_shard(rank, world)
        rd.preload()

        def rstep():
            sts = []
            for fn in fns:
                if fn == "introns":
                    rd.count_introns()
                else:
                    rd.count_reads_raw({"unstranded": 0, "stranded": 1, "weighted": 2}[fn], intervals, gene_intervals)
                sts.append(rd.stats())
            return None, sts

        r_total, res_sts, _, _ = timed(rstep, args.steps, 2)
        rd.drop_preload()
        resident = {"value": n_records * args.steps / r_total, "unit": UNIT, "ms_per_step": 1000 * r_total / args.steps,
                    "note": "compressed file already in HBM (mbfbam_preload), dense results only: kernels, not the metric"}

    # ---- stage times and rooflines (rank 0's shard; CUDA events on the launching streams)
    src_sts = [s for step_sts in (res_sts or e2e_sts) for s in step_sts]
    n_calls = max(1, len(src_sts))
    mean = lambda k: sum(s[k] for s in src_sts) / n_calls  # noqa: E731
    peak, peak_src = measured_peak_gbs()
    C, U, T, R, W = mean("compressed_bytes"), mean("uncompressed_bytes"), mean("n_tokens"), mean("n_records"), max(1.0, mean("n_windows"))
    traffic = ncu_traffic()

    def roof(kernel, ms, alg_bytes, note):
        ach = alg_bytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        tr = traffic.get(kernel)
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                "traffic": int(tr["dram_bytes_per_algorithmic_byte"] * alg_bytes / W) if tr else None,
                "traffic_source": tr.get("source") if tr else None,
                # the same kernel alone on the GPU (ncu, one launch): the CUDA-event spans above contain the time a
                # launch waits for SMs held by the kernels of neighbouring windows, which dominates for the short kernels
                "achieved_alone": tr.get("achieved_alone_gbs") if tr else None,
                "frac_alone": (tr["achieved_alone_gbs"] / peak) if tr and tr.get("achieved_alone_gbs") and peak else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes / W), "launches_per_call": int(W),
                "ms_per_call": ms, "note": note}

    ms_tok, ms_inf = mean("ms_gpu_tokens"), mean("ms_gpu_inflate")
    roofs = [
        roof("k_tokens", ms_tok, C + 4 * T, "K1 pass 1 (Huffman decode): compressed stream read once + one 4-byte token per DEFLATE symbol"),
        roof("k_lz_resolve", max(0.0, ms_inf - ms_tok), 4 * T + U, "K1 pass 2 (LZ77): tokens read once + inflated bytes written once"),
        roof("k_scan", mean("ms_gpu_scan"), 2 * C, "K0: every compressed byte tested twice (count, emit)"),
        roof("k_frame", mean("ms_gpu_frame"), 8 * R, "K2: block_size of every record read, its offset written"),
        roof("k_count", mean("ms_gpu_count"), 96 * R, "K3-K5: ~96 bytes of header/CIGAR/aux per record"),
    ]
    top = dict(max(roofs, key=lambda r: r["ms_per_call"]))
    top["k1_both_passes"] = {"kernels": "k_tokens + k_lz_resolve", "algorithmic_bytes_per_call": int(C + U), "ms_per_call": ms_inf,
                             "achieved": (C + U) / (ms_inf / 1e3) / 1e9 if ms_inf > 0 else 0.0, "unit": "GB/s",
                             "note": "SURVEY 8(d) counts K1 as one kernel with C + U algorithmic bytes"}

    # ---- measured pinned H2D rate of this box (the ingest roofline's denominator): every rank at once (on a
    #      multi-GPU box the GPUs share the host's PCIe / IOMMU path: the aggregate is what bounds N > 1), then rank 0
    #      alone
    def h2d_rate(sync_ranks):
        import torch

        hb = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        db = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:%d" % local)
        db.copy_(hb, non_blocking=True)
        torch.cuda.synchronize()
        if sync_ranks:
            barrier()
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            db.copy_(hb, non_blocking=True)
            e1.record()
            e1.synchronize()
            best = max(best, (1 << 30) / (e0.elapsed_time(e1) / 1e3) / 1e9)
        return best

    h2d_peak = h2d_peak_all = None
    try:
        if dist is not None:
            h2d_peak_all = greduce(h2d_rate(True))
            barrier()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("concurrent h2d measurement failed: %s\n" % (e,))

    if rank != 0:
        if dist is not None:
            dist.barrier()  # rank 0 measures its solo H2D rate while the others wait here
            dist.destroy_process_group()
        return 0
    try:
        h2d_peak = h2d_rate(False)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("h2d peak measurement failed: %s\n" % (e,))
    if dist is not None:
        dist.barrier()
    if h2d_peak_all is None and h2d_peak is not None:
        h2d_peak_all = h2d_peak * args.gpus

    # ---- parity at the benchmarked scale: the C oracle over the WHOLE file, every key of every result
    parity = {"checked": False}
    if not args.no_parity:
        from oracle import c_oracle

        t0 = time.time()
        nkeys = 0
        for fn in fns:
            got = e2e_last[fn]
            if fn == "introns":
                want = c_oracle.count_introns(path, nthreads=threads)
                assert got == want, "count_introns differs from the oracle"
                nkeys += sum(len(d) for d in got.values())
            elif fn == "weighted":
                want = c_oracle.count_reads_weighted(path, None, intervals, gene_intervals, nthreads=threads)
                for a, b in zip(got, want):
                    assert set(a) == set(b), "count_reads_weighted: key sets differ"
                    for k in a:
                        assert abs(a[k] - b[k]) <= 1e-9 * max(1.0, abs(b[k])), ("count_reads_weighted", k, a[k], b[k])
                    nkeys += len(a)
            elif fn == "stranded":
                want = c_oracle.count_reads_stranded(path, None, intervals, gene_intervals, nthreads=threads)
                assert tuple(got) == tuple(want), "count_reads_stranded differs from the oracle"
                nkeys += len(got[0]) + len(got[1])
            else:
                want = c_oracle.count_reads_unstranded(path, None, intervals, gene_intervals, nthreads=threads)
                assert got == want, "count_reads_unstranded differs from the oracle"
                nkeys += len(got)
        parity = {"checked": True, "oracle": "oracle/mbf_oracle.c over the whole file, %d threads" % threads, "functions": fns,
                  "keys_compared": nkeys, "oracle_s": round(time.time() - t0, 2), "result": "bit-exact (weighted: 1e-9 relative)"}

    # ---- CPU baseline on this box's cores, bounded sample
    from oracle import c_oracle

    n_chunks = len(c_oracle.chunks(path, None, gene_intervals))
    cpu_v, cpu_nrec, cpu_nck, cpu_dt = cpu_sample(path, intervals, gene_intervals, fns, threads, max(16, n_chunks // 4))

    e2e = n_records * args.steps / e2e_total
    ingest_gbs = h2d / (e2e_total / args.steps) / 1e9
    api = ("mbf_bam_b200." + "+".join("count_introns" if f == "introns" else "count_reads_" + f for f in fns)) if dist is None \
        else "mbf_bam_b200.dist (one rank per GPU, NCCL all-reduce): " + "+".join(fns)
    line = {
        "metric": METRIC, "value": e2e, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * e2e_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000 * e2e_total / args.steps,
                "includes": "host file (page cache) -> H2D -> kernels -> D2H -> merge -> Python dicts, through " + api,
                "h2d_from_page_locked_file_bytes_per_step": int(h2d_pinned)},
        "e2e_first_calls": {"ms": [round(1000 * t, 1) for t in t_first],
                            "note": "the first two calls of the process on this file, before anything is page-locked (copy through "
                                    "pinned staging buffers); MBF_BAM_PIN_AFTER (default 3) decides when the mapping is page-locked"},
        "resident": resident,
        "gpu_launches": launches,
        "roofline": top,
        "roofline_all": roofs,
        "ingest": {"h2d_gbs": ingest_gbs, "host_read_gbs": ingest_gbs, "peak_gbs_one_gpu": h2d_peak, "peak_gbs_all_gpus_at_once": h2d_peak_all,
                   "frac": ingest_gbs / h2d_peak_all if h2d_peak_all else None,
                   "frac_of_n_times_one_gpu": ingest_gbs / (h2d_peak * args.gpus) if h2d_peak else None,
                   "note": "compressed bytes over PCIe per second of e2e (all GPUs) against the pinned H2D rate measured in this run "
                           "with every GPU copying at once (the GPUs of a box share the host side of PCIe)"},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d chunks (%d reads, %.1f s wall), C restatement of the reference (zlib + pthreads)" % (
                             cpu_nck, n_chunks, cpu_nrec, cpu_dt)},
        "parity": parity,
        "clocks": clocks,
        "shard_imbalance": imbalance,
        "stage_ms_rank0": {k: mean(k) for k in ("ms_gpu_scan", "ms_gpu_inflate", "ms_gpu_tokens", "ms_gpu_frame", "ms_gpu_count",
                                                "ms_gpu_all", "ms_total", "ms_pin")},
        "bam_generation_s": round(t_gen, 1),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
