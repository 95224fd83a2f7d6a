#!/usr/bin/env python
"""bench.py -- aligned reads/sec counted from a BAM, BASELINE.json configs[1] shape:
count_reads_stranded on synthetic 100 bp reads, 25 human-like contigs, 20k genes / ~60k exon
intervals, 50M reads per GPU (weak scaling: N GPUs -> one N*50M-read file sharded by chunk range).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R_per_gpu]

One "step" = one complete count_reads_stranded job over the whole file.
  value : reads/s with the compressed BAM already resident in HBM (mbfbam_preload) -> kernels ->
          dense counts -> (NCCL all-reduce when N > 1).
  e2e   : reads/s through the public API from HOST memory (page cache): pread -> pinned -> H2D ->
          kernels -> D2H -> Python dicts.
  roofline : k_tokens_bf (pass 1 of the two-pass inflate, the longest kernel): algorithmic bytes
          (compressed stream read once + one 4-byte token written per DEFLATE symbol) over its CUDA-event
          time on the launching stream, against the measured HBM copy bandwidth.
  cpu_baseline : the CPU restatement of the reference (oracle/, zlib + pthreads, one task per chunk like
          rayon in counters.rs:217-259) on all host cores over a bounded sample of the same file.
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
# dram__bytes_read.sum + dram__bytes_write.sum of k_tokens_bf<8> divided by its algorithmic bytes (C + 4 T),
# from the ncu --set full capture summarised in profiles/r01tp_ncu_full_summary.txt
K1_TRAFFIC_RATIO_NCU = 0.99
K1_NCU_PROFILE = "profiles/r01tp_ncu_full_summary.txt"


def env_int(name, default):
    v = os.environ.get(name)
    return int(v) if v else default


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def dist_env():
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    return rank, world, local


def ensure_bam(workdir, reads_total, rank, world, barrier):
    from mbf_bam_b200.synth import model

    c = model.CONFIGS["cfg2"]
    intervals, gene_intervals, genes = model.make_gene_model(c["contigs"], c["n_genes"], c["seed"], gene_mode=False,
                                                             span=c["span"], mean_exons=4.6)
    level = env_int("MBF_BENCH_LEVEL", 6)
    name = "cfg2_%dM_l%d.bam" % (reads_total // 1_000_000, level)

    def complete(p):
        return os.path.exists(p) and os.path.exists(p + ".bai") and os.path.exists(p + ".ok")

    # Where the synthetic file lives: the given directory, else the first of /tmp, /dev/shm, the repo with room for
    # it (N x 3.2 GB at N GPUs).  Rank 0 decides and leaves a note; the other ranks read it after the barrier.
    note = os.path.join("/tmp", "mbf_bench_%s.where" % name)
    if rank == 0:
        cands = [workdir] if workdir else ["/tmp/mbf_bench", "/dev/shm/mbf_bench", os.path.join(ROOT, ".bench")]
        need = int(reads_total * 70 * 1.1)
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
                                  in_gene=c["in_gene"], spliced=c["spliced"], mm=c["mm"])
        t_gen = time.time() - t0
        open(path + ".ok", "w").write(json.dumps(info))
    barrier()
    if rank != 0:
        path = os.path.join(open(note).read().strip(), name)
    info = json.load(open(path + ".ok"))
    return path, intervals, gene_intervals, info, t_gen


def cpu_sample(path, intervals, gene_intervals, n_chunks_total, threads, target_chunks):
    from oracle import c_oracle

    t0 = time.time()
    r = c_oracle.raw_counts(path, None, intervals, gene_intervals, 1, nthreads=threads, sample_chunks=target_chunks)
    dt = time.time() - t0
    return r["n_records"] / dt, r["n_records"], r["n_chunks"], dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=env_int("MBF_BENCH_READS", 50_000_000), help="reads per GPU")
    ap.add_argument("--workdir", default=os.environ.get("MBF_BENCH_DIR"),
                    help="directory of the synthetic BAM (default: /tmp, /dev/shm or the repo, whichever has room)")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    threads = os.cpu_count() or 1
    # one rank per GPU shares the host cores: split the file-staging threads between the ranks
    # host threads copying file bytes into pinned memory: 3/4 of this rank's share of the cores (the issuing thread,
    # the staging helper and the driver's own threads need the rest; measured 12 > 16 > 8 on a 16-core box)
    os.environ.setdefault("MBF_BAM_READ_THREADS", str(max(2, min(16, (threads // max(1, world)) * 3 // 4))))

    if args.impl == "reference" and rank != 0:
        return 0

    dist = None
    if world > 1 and args.impl == "ours":
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()

    reads_total = args.reads * max(1, args.gpus)
    import __graft_entry__ as g

    if rank == 0:
        g.build()
    barrier()
    path, intervals, gene_intervals, info, t_gen = ensure_bam(args.workdir, reads_total, rank, world, barrier)
    n_records = int(info["records"])
    n_intervals = sum(len(x[2]) for v in intervals.values() for x in v)
    config = {"workload": "configs[1]: count_reads_stranded, synthetic 100bp single-end sorted+indexed BAM, 25 human-like contigs, "
                          "%d genes / %d exon intervals, zlib level %d" % (sum(len(v) for v in intervals.values()), n_intervals,
                                                                          env_int("MBF_BENCH_LEVEL", 6)),
              "reads": n_records, "reads_per_gpu": args.reads, "bam_bytes": int(info["bytes"]),
              "l2": "inputs larger than L2 (compressed file %.1f GB, inflated %.1f GB per pass)" % (
                  info["bytes"] / 1e9, n_records * 207 / 1e9),
              "sharding": "chunk-range shards of one file, one rank per GPU" if args.gpus > 1 else "single shard"}

    if args.impl == "reference":
        # the reference's CPU path cannot be built here (no cargo/rustc): time its CPU restatement (oracle/)
        from oracle import c_oracle

        n_chunks = len(c_oracle.chunks(path, None, gene_intervals))
        target = max(16, n_chunks // 6)
        vals = []
        for i in range(args.warmup + args.steps):
            v, nrec, nck, dt = cpu_sample(path, intervals, gene_intervals, n_chunks, threads, target)
            if i >= args.warmup:
                vals.append((v, dt))
        value = statistics.mean(v for v, _ in vals)
        sample = "%d of %d chunks (%d reads) per step, %d threads" % (nck, n_chunks, nrec, threads)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000 * statistics.mean(dt for _, dt in vals), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np

    import mbf_bam_b200 as mb

    if mb.device_count() < 1:
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    rd = mb.Reader(path)
    rd.set_device(local)
    rd.set_shard(rank, world)

    def merge(raw):
        """dense per-shard results -> whole-job results (NCCL sum when the file is sharded)"""
        if dist is None:
            return raw
        import torch

        fwd = torch.frombuffer(bytearray(raw[0]), dtype=torch.int64).cuda()
        rev = torch.frombuffer(bytearray(raw[1]), dtype=torch.int64).cuda()
        extra = torch.tensor([raw[5]], dtype=torch.int64).cuda()
        sc = torch.frombuffer(bytearray(raw[4]), dtype=torch.uint8).to(torch.int32).cuda()
        for t in (fwd, rev, extra, sc):
            dist.all_reduce(t)
        return (fwd.cpu().numpy().tobytes(), rev.cpu().numpy().tobytes(), raw[2], raw[3],
                (sc > 0).to(torch.uint8).cpu().numpy().tobytes(), int(extra.item()))

    def step(resident):
        t0 = time.perf_counter()
        if dist is None and not resident and not pinned_mode[0]:
            # the call a user makes: the module-level drop-in function (opens BAM + BAI, counts, builds the dicts)
            res = mb.count_reads_stranded(path, None, intervals, gene_intervals)
            return time.perf_counter() - t0, mb.last_stats(), None, res
        raw = rd.count_reads_raw(1, intervals, gene_intervals)
        st = rd.stats()
        raw = merge(raw)
        res = rd.assemble(1, intervals, *raw) if (rank == 0 and not resident) else None
        return time.perf_counter() - t0, st, raw, res

    pinned_mode = [False]

    def timed(resident, steps, warmup):
        for _ in range(warmup):
            step(resident)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        sts = []
        last = None
        for _ in range(steps):
            dt, st, raw, res = step(resident)
            sts.append(st)
            last = (raw, res)
        barrier()
        total = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None
        if dist is not None:
            import torch

            t = torch.tensor([total], dtype=torch.float64).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, sts, last, clocks

    # ---- e2e: host file -> dicts
    e2e_total, e2e_sts, e2e_last, _ = timed(False, args.steps, args.warmup)
    # ---- same with the file mapping page-locked once per reader (mbfbam_pin_file): extra information only
    pin_total = None
    try:
        # single-GPU runs only: with N ranks every rank would page-lock the whole N x 3 GB file for a figure that
        # is extra information
        if dist is None and rd.pin_file():
            pinned_mode[0] = True
            pin_total, _, pin_last, _ = timed(False, args.steps, 1)
            pinned_mode[0] = False
            rd.unpin_file()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("pin_file failed: %s\n" % (e,))
        pinned_mode[0] = False
    # ---- value: compressed bytes resident in HBM
    rd.preload()
    val_total, val_sts, val_last, clocks = timed(True, args.steps, args.warmup)
    rd.drop_preload()

    def gsum(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64).cuda()
        dist.all_reduce(t)
        return float(t.item())

    launches = int(gsum(sum(s["n_kernel_launches"] for s in val_sts)))
    h2d = gsum(statistics.mean(s["h2d_bytes"] for s in e2e_sts))
    d2h = gsum(statistics.mean(s["d2h_bytes"] for s in e2e_sts))
    # roofline of the dominant kernel on rank 0's shard
    st0 = val_sts[-1]
    ms_infl = statistics.mean(s["ms_gpu_tokens"] for s in val_sts)
    alg_bytes = st0["compressed_bytes"] + 4 * st0["n_tokens"]
    ms_k1 = statistics.mean(s["ms_gpu_inflate"] for s in val_sts)
    alg_cu = st0["compressed_bytes"] + st0["uncompressed_bytes"]
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (ms_infl / 1e3) / 1e9 if ms_infl > 0 else 0.0
    stage_ms = {k: statistics.mean(s[k] for s in val_sts) for k in
                ("ms_gpu_scan", "ms_gpu_inflate", "ms_gpu_tokens", "ms_gpu_frame", "ms_gpu_count", "ms_gpu_all", "ms_total")}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # parity spot check inside the bench: whole-job totals of both arms must agree
    fwd_total = int(np.frombuffer(val_last[0][0], dtype=np.uint64).sum())
    rev_total = int(np.frombuffer(val_last[0][1], dtype=np.uint64).sum())
    assert e2e_last[1][0]["_total"] == fwd_total and e2e_last[1][1]["_total"] == rev_total, "resident and streamed results differ"
    assert e2e_last[1][0]["_outside"] == int(val_last[0][5])

    # ---- CPU baseline on this box's cores, bounded sample
    from oracle import c_oracle

    n_chunks = len(c_oracle.chunks(path, None, gene_intervals))
    cpu_v, cpu_nrec, cpu_nck, cpu_dt = cpu_sample(path, intervals, gene_intervals, n_chunks, threads, max(16, n_chunks // 4))

    value = n_records * args.steps / val_total
    e2e = n_records * args.steps / e2e_total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * val_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000 * e2e_total / args.steps, "includes": "pread from page cache, H2D, kernels, D2H, dict assembly"},
        "e2e_pinned_file": None if pin_total is None else {
            "value": n_records * args.steps / pin_total, "unit": UNIT, "ms_per_step": 1000 * pin_total / args.steps,
            "note": "file mapping page-locked once per reader (mbfbam_pin_file); H2D straight from the page cache"},
        "gpu_launches": launches,
        "roofline": {"kernel": "k_tokens_bf<8>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": int(alg_bytes * K1_TRAFFIC_RATIO_NCU / max(1, st0["n_windows"])),
                     "traffic_note": "per launch; dram bytes / algorithmic bytes = %.2f measured with ncu (%s)" % (
                         K1_TRAFFIC_RATIO_NCU, K1_NCU_PROFILE),
                     "peak_source": peak_src, "algorithmic_bytes_per_step": int(alg_bytes),
                     "algorithmic_bytes_per_launch": int(alg_bytes / max(1, st0["n_windows"])),
                     "launches_per_step": int(st0["n_windows"]), "ms_per_step": ms_infl,
                     # SURVEY 8(d) counts K1 as one kernel with C + U algorithmic bytes: the same figure for the pair
                     "k1_both_passes": {"kernels": "k_tokens_bf + k_lz_resolve", "algorithmic_bytes_per_step": int(alg_cu),
                                        "ms_per_step": ms_k1, "achieved": alg_cu / (ms_k1 / 1e3) / 1e9 if ms_k1 > 0 else 0.0,
                                        "unit": "GB/s"},
                     "note": "DEFLATE symbol decode is a dependency chain: the kernel is latency/issue bound, not HBM bound "
                             "(DESIGN.md, K1 roofline honesty); frac is small by construction"},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d chunks (%d reads, %.1f s wall), C restatement of the reference (zlib + pthreads)" % (
                             cpu_nck, n_chunks, cpu_nrec, cpu_dt)},
        "clocks": clocks,
        "stage_ms_rank0": stage_ms,
        "totals": {"forward": fwd_total, "reverse": rev_total, "outside": int(val_last[0][5]), "bam_generation_s": round(t_gen, 1)},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
