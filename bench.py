#!/usr/bin/env python
"""Benchmark of the BWA-MEM hot path (BASELINE.json metric: reads/s, 2x150 bp paired end).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--pairs P] [--ref-len L]

One "step" is one pass of the device pipeline over the whole synthetic job (P read pairs cut into
-K 10,000,000-base batches exactly as bseq_read does).  Workload at N=1: BASELINE.json configs[1]
(1 M x 2x150 bp pairs vs a 100 Mbp synthetic reference; both scale down with --pairs/--ref-len).
Reported on one JSON line:
  value      reads/s with the reads already resident in HBM (kernels + the small per-batch
             host<->device control traffic; SAM text stays on the device)
  e2e        reads/s through the reference-facing call b200bwa_main_to (FASTQ files -> SAM file, the
             argv SparkBWA builds), host parsing, H2D, kernels, D2H and file write inside the timed region
  roofline   seeding kernel: algorithmic bytes (64 B per occ-block touch, counted by the kernel the way
             SURVEY.md 8(d) defines them) / its CUDA-event duration, against the measured HBM peak
  sw         DP work: cells of ksw_extend2 + ksw_global2 + local SW over the DP stages' device time
  cpu_baseline  the reference's own `bwa mem -t <host cores>` (oracle/_ref) on a bounded sample
Multi-GPU (torchrun, one rank per GPU): STRONG scaling of the same job at every N.  `value`: the job's -K
batches are dealt over the ranks (the K passes of the timed region are queued back to back, pass p batch b -> rank
(p * n_batches + b) mod N, each with its absolute n_processed; no drain between passes, like the steps of a training
loop), the index image is broadcast once over NCCL, no collective on the data path; time = max over ranks.  `e2e`: ONE call of the product entry
point on rank 0 drives all N GPUs (one reader, N engines, one ordered SAM file -- what a Spark executor's JNI call
does); the other ranks wait on the rendezvous store without touching their GPUs.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
LIB = os.environ.get("B200BWA_LIB", os.path.join(ROOT, "sparkbwa_b200", "libbwa.so"))
ORACLE = os.path.join(ROOT, "oracle", "_ref", "bwa")


class Times(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("h2d", "seed", "sa", "chain", "extend", "pestat", "finalize", "gather", "d2h", "total")] + \
               [("launches", ctypes.c_int), ("n_seed_tasks", ctypes.c_int64), ("n_regs", ctypes.c_int64), ("sam_bytes", ctypes.c_int64),
                ("rescue", ctypes.c_float), ("n_rescue_tasks", ctypes.c_int64),
                ("k_xext", ctypes.c_float), ("k_resc_sw", ctypes.c_float), ("k_galn", ctypes.c_float)]


def load_lib():
    lib = ctypes.CDLL(LIB)
    lib.b200bwa_last_error.restype = ctypes.c_char_p
    lib.b200bwa_index_load.restype = ctypes.c_void_p
    lib.b200bwa_index_load.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.b200bwa_index_adopt.restype = ctypes.c_void_p
    lib.b200bwa_index_adopt.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.b200bwa_set_options.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.b200bwa_set_paired.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.b200bwa_upload_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p,
                                         ctypes.c_char_p, ctypes.c_int64]
    lib.b200bwa_run_resident_range.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Times)]
    lib.b200bwa_counters.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.b200bwa_process_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.b200bwa_main_to.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p]
    lib.b200bwa_index_free.argtypes = [ctypes.c_void_p]
    return lib


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_ev = gpu, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_ev.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5)
                self.rows.append([x.strip() for x in r.stdout.decode().strip().split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workdir(args):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    d = os.path.join(base, f"b2bench_{args.workload}_L{args.ref_len}_P{args.pairs}_s{args.seed}")
    os.makedirs(d, exist_ok=True)
    return d


def prepare_reference(args, d, rank):
    """Reference FASTA-less: contigs in memory -> index files written by sparkbwa_b200.index (rank 0)."""
    from sparkbwa_b200 import synth, index
    prefix = os.path.join(d, "ref")
    n_ctg = 4 if args.ref_len >= 20_000_000 else 2
    ctg = synth.make_reference(args.ref_len, n_contigs=n_ctg, seed=args.seed, repeat_frac=0.02)
    if rank == 0 and not os.path.exists(prefix + ".sa"):
        t = time.time()
        index.build_index(ctg, prefix + ".tmp")
        for e in ("pac", "ann", "amb", "bwt", "sa"):
            os.replace(prefix + ".tmp." + e, prefix + "." + e)
        sys.stderr.write(f"[bench] index built in {time.time() - t:.1f}s\n")
    return ctg, prefix


def prepare_reads(args, d, ctg, rank):
    from sparkbwa_b200 import synth
    f1 = os.path.join(d, "r_1.fq")
    f2 = os.path.join(d, "r_2.fq")
    s1, s2 = [], []
    t = time.time()
    if args.paired:
        gen = synth.simulate_pairs_fast(ctg, args.pairs, 150, seed=args.seed * 1000 + 17, sub=0.01, indel=0.001)
    else:  # single-end 250 bp: mate 1 of the same simulator, up to six single-base indels per read
        gen = synth.simulate_pairs_fast(ctg, args.pairs, 250, seed=args.seed * 1000 + 19, sub=0.035, indel=0.015, indel_rounds=6,
                                        ins_mean=600.0, ins_sd=60.0)
    for a, b in gen:
        s1.append(a); s2.append(b)
    if rank == 0 and not os.path.exists(f2 if args.paired else f1):
        synth.write_fastq_fixed(f1 + ".tmp", s1, 0, b"/1" if args.paired else b""); os.replace(f1 + ".tmp", f1)
        if args.paired:
            synth.write_fastq_fixed(f2 + ".tmp", s2, 0, b"/2"); os.replace(f2 + ".tmp", f2)
    sys.stderr.write(f"[bench] rank {rank}: {args.pairs} {'pairs' if args.paired else 'reads'} simulated in {time.time() - t:.1f}s\n")
    return np.concatenate(s1), (np.concatenate(s2) if args.paired else None), f1, (f2 if args.paired else None)


def cut_batches(n_reads, read_len, chunk):
    """bseq_read (bwa.c:42-75): a batch ends once its bases reach `chunk` and the read count is even (reads are taken
    two at a time from a pair of files, one at a time from a single file)."""
    out, first = [], 0
    per = read_len
    while first < n_reads:
        n = 0
        size = 0
        while first + n < n_reads:
            n += 1; size += per
            if size >= chunk and n % 2 == 0:
                break
        out.append((first, n))
        first += n
    return out


def cpu_reference(args, prefix, f1, f2, sample_pairs, threads):
    """Times the reference's own bwa mem on the host cores over the first `sample_pairs` pairs."""
    d = os.path.dirname(f1)
    files = [(f1, os.path.join(d, "cpu_1.fq"))] + ([(f2, os.path.join(d, "cpu_2.fq"))] if f2 else [])
    for src, dst in files:
        with open(src, "rb") as fi, open(dst, "wb") as fo:
            l2 = sum(len(fi.readline()) for _ in range(4))   # fixed-width records: size of one record = 4 lines
            fi.seek(0)
            fo.write(fi.read(l2 * sample_pairs))
    # index load warm-up is part of the reference run (as it is for a SparkBWA task); report wall time of the whole call
    t = time.time()
    r = subprocess.run([ORACLE, "mem", "-t", str(threads), "-K", str(args.K), prefix] + [x[1] for x in files], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE)
    dt = time.time() - t
    if r.returncode != 0:
        raise RuntimeError("oracle bwa failed: " + r.stderr.decode()[-500:])
    # subtract nothing: honest wall clock; parse bwa's own per-batch times for the note
    proc = [float(l.split(" in ")[1].split(" CPU")[0]) for l in r.stderr.decode().splitlines() if "Processed" in l]
    return (2 if f2 else 1) * sample_pairs / dt, dt, sum(proc)


def main():
    # the driver reads ONE JSON line from stdout: keep everything else (NCCL banners, library chatter) on stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--workload", default="pe150", choices=["pe150", "se250"],
                    help="pe150: BASELINE.json configs[1]/[4] (2x150 bp pairs, 1 %% error); se250: configs[3] (single-end 250 bp, 3.5 %% "
                         "substitutions + 1.5 %% indels: wide ksw_extend2 bands, band retries)")
    ap.add_argument("--pairs", type=int, default=0, help="pairs (pe150; default 1,000,000) or reads (se250; default 400,000)")
    ap.add_argument("--ref-len", type=int, default=100_000_000)
    ap.add_argument("--K", type=int, default=10_000_000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--cpu-sample-pairs", type=int, default=0,
                    help="pairs timed by the CPU legs (0: the whole workload for --impl reference, 200000 for the engine arm's cpu_baseline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e-large", action="store_true", help="skip the N x pairs companion job of the multi-GPU e2e leg")
    args = ap.parse_args()
    args.paired = args.workload == "pe150"
    args.read_len = 150 if args.paired else 250
    if args.pairs <= 0:
        args.pairs = 1_000_000 if args.paired else 400_000

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host_cores = os.cpu_count() or 1
    if args.paired:
        workload = f"BWA-MEM paired-end {args.pairs} x 2x150bp vs {args.ref_len} bp synthetic reference (2% planted repeats), -K {args.K}"
        metric = "BWA-MEM reads/sec 2x150bp PE"
    else:
        workload = (f"BWA-MEM single-end {args.pairs} x 250bp at 3.5% substitutions + 1.5% indels vs {args.ref_len} bp synthetic reference "
                    f"(2% planted repeats), -K {args.K}")
        metric = "BWA-MEM reads/sec 250bp SE"

    if args.impl == "reference":
        if rank != 0:
            return
        d = workdir(args)
        try:  # data preparation only (outside the timed path): the GPU index builder writes files byte-identical to `bwa index`
            import torch
            have_cuda = torch.cuda.is_available()
        except Exception:
            have_cuda = False
        ctg, prefix = prepare_reference(args, d, 0) if have_cuda else prepare_reference_cpu_only(args, d)
        s1, s2, f1, f2 = prepare_reads(args, d, ctg, 0)
        sample = min(args.pairs, args.cpu_sample_pairs or args.pairs)
        # index load alone (the same call on a single pair): reported, not subtracted
        _, t_index, _ = cpu_reference(args, prefix, f1, f2, 1, host_cores)
        vals = []
        for i in range(args.warmup + args.steps):
            v, dt, _ = cpu_reference(args, prefix, f1, f2, sample, host_cores)
            if i >= args.warmup:
                vals.append((v, dt))
        v = float(np.mean([x[0] for x in vals]))
        ms = float(np.mean([x[1] for x in vals])) * 1e3
        print(file=real_stdout, end="")
        real_stdout.write(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "timing": "wall clock of the whole `bwa mem` call incl. its index load from the page cache "
                       f"({t_index:.2f} s of the {ms / 1e3:.2f} s; the engine arm's e2e calls find the index cached in HBM, as "
                       "successive tasks of one executor would)"},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": host_cores, "kind": "reference", "index_load_s": t_index,
                             "sample": ("the whole workload" if sample == args.pairs else f"first {sample} pairs of the workload") +
                                       f", bwa mem -t {host_cores} -K {args.K}"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}) + "\n")
        real_stdout.flush()
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_lib()
    d = workdir(args)
    ctg, prefix = prepare_reference(args, d, rank)
    if world > 1:
        dist.barrier()
    s1, s2, f1, f2 = prepare_reads(args, d, ctg, rank)
    del ctg

    # ---- index image: read by rank 0, broadcast over NCCL, adopted in place by every rank ----------
    def file_tensor(path, skip, pad):
        a = np.fromfile(path, dtype=np.uint8)[skip:]
        t = torch.zeros(a.size + pad, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t[: a.size] = torch.from_numpy(a).cuda()
        return t
    bwt_t = file_tensor(prefix + ".bwt", 40, 256)
    sa_raw = np.fromfile(prefix + ".sa", dtype=np.uint8)[56 - 8:].copy()  # keep one slot for sa[0] = -1
    sa_raw[:8] = 255
    sa_t = torch.zeros(sa_raw.size + 64, dtype=torch.uint8, device="cuda")
    if rank == 0:
        sa_t[: sa_raw.size] = torch.from_numpy(sa_raw).cuda()
    pac_t = file_tensor(prefix + ".pac", 0, 64)
    if world > 1:
        for t in (bwt_t, sa_t, pac_t):
            dist.broadcast(t, src=0)
    torch.cuda.synchronize()
    ix = lib.b200bwa_index_adopt(prefix.encode(), local, bwt_t.data_ptr(), sa_t.data_ptr(), pac_t.data_ptr())
    if not ix:
        raise SystemExit("index adopt failed: " + lib.b200bwa_last_error().decode())
    assert lib.b200bwa_set_options(ix, f"-K {args.K}".encode()) == 0
    lib.b200bwa_set_paired(ix, 1 if args.paired else 0)

    # ---- resident job ------------------------------------------------------------------------------
    RL = args.read_len
    if args.paired:
        n_reads = 2 * args.pairs
        inter = np.empty((n_reads, RL), dtype=np.uint8)
        inter[0::2] = s1; inter[1::2] = s2
        names = b"".join(b"r%09d\0r%09d\0" % (i, i) for i in range(args.pairs))
    else:
        n_reads = args.pairs
        inter = s1
        names = b"".join(b"r%09d\0" % i for i in range(args.pairs))
    del s1, s2
    seq = inter.tobytes()
    qual = b"I" * len(seq)
    off = (np.arange(n_reads + 1, dtype=np.int64) * RL)
    if lib.b200bwa_upload_batch(ix, n_reads, seq, qual, off.ctypes.data, names, 0) != 0:
        raise SystemExit("upload failed: " + lib.b200bwa_last_error().decode())
    del inter
    all_batches = cut_batches(n_reads, RL, args.K)
    # strong scaling: the passes of one measurement are queued back to back (like the steps of a training loop: no drain
    # between them), batch b of pass p -> rank (p * n_batches + b) mod N, each with its absolute n_processed
    deal_rank, deal_world = rank, world
    if os.environ.get("B200BWA_BENCH_DEAL"):   # "r/w": run, on this one GPU, the share rank r of w ranks would get (debugging aid)
        deal_rank, deal_world = (int(x) for x in os.environ["B200BWA_BENCH_DEAL"].split("/"))

    def dealt(n_pass):
        seq = [bt for _ in range(n_pass) for bt in all_batches]
        return seq[deal_rank::deal_world]
    batches = all_batches[rank::world]
    stagger_s = float(os.environ.get("B200BWA_BENCH_STAGGER_MS", "0")) / 1e3
    lib.b200bwa_n_workers.argtypes = [ctypes.c_void_p]
    n_workers = max(1, min(args.workers, int(lib.b200bwa_n_workers(ix))))

    def one_pass(collect=None, batches=batches):
        lock = threading.Lock()
        nxt = [0]
        err = []

        def work(wi):
            t = Times()
            if stagger_s > 0:
                time.sleep(wi * stagger_s)
            while True:
                with lock:
                    b = nxt[0]; nxt[0] += 1
                if b >= len(batches) or err:
                    return
                first, cnt = batches[b]
                rc = lib.b200bwa_run_resident_range(ix, wi, first, cnt, 1 if args.paired else 0, 0, None, None, ctypes.byref(t))
                if rc != 0:
                    err.append(lib.b200bwa_last_error().decode()); return
                if collect is not None:
                    collect.append({k: getattr(t, k) for k, _ in Times._fields_})
        th = [threading.Thread(target=work, args=(w,)) for w in range(n_workers)]
        [x.start() for x in th]; [x.join() for x in th]
        if err:
            raise SystemExit("pipeline failed: " + err[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_pass(batches=dealt(args.warmup))
    cnt0 = (ctypes.c_ulonglong * 16)()
    for w in range(n_workers):
        lib.b200bwa_counters(ix, w, cnt0, 1)
    pc = (ctypes.c_ulonglong * 3)()
    lib.b200bwa_process_counters(pc, 1)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trace = [] if os.environ.get("B200BWA_BENCH_TRACE") else None
    one_pass(trace, batches=dealt(args.steps))
    torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    ms_total = e0.elapsed_time(e1)
    sampler.stop_ev.set(); sampler.join()
    if trace:   # per-batch stage times of the concurrent passes (each batch's own CUDA events): mean / max over the batches
        for k in ("seed", "sa", "chain", "extend", "pestat", "rescue", "finalize", "gather", "total"):
            v = [t[k] for t in trace]
            sys.stderr.write(f"[bench trace] {k:9s} mean {np.mean(v):8.2f} ms  max {np.max(v):8.2f} ms  ({len(v)} batches)\n")
    lib.b200bwa_process_counters(pc, 0)
    launches = int(pc[2])
    tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    ms_total = float(tt.item())
    launches = int(tl.item())
    ms_step = ms_total / args.steps
    value = n_reads / (ms_step / 1e3)   # the whole job's reads over the slowest rank's time

    # ---- single-stream pass for clean per-kernel CUDA-event durations + algorithmic counters ------
    for w in range(n_workers):
        lib.b200bwa_counters(ix, w, cnt0, 1)
    n_workers_saved, n_workers = n_workers, 1
    stage = []
    batches = all_batches if rank == 0 else []   # per-kernel figures: rank 0 runs every batch of the job on one stream
    one_pass(stage, batches)
    n_workers = n_workers_saved
    cnt = (ctypes.c_ulonglong * 16)()
    lib.b200bwa_counters(ix, 0, cnt, 1)
    if rank != 0:
        stage = [{k: 0.0 for k, _ in Times._fields_}]
    tot = {k: float(sum(s[k] for s in stage)) for k in ("seed", "sa", "chain", "extend", "pestat", "rescue", "finalize", "gather", "total",
                                                        "k_xext", "k_resc_sw", "k_galn")}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    seed_bytes = 64.0 * cnt[0]
    seed_gbs = seed_bytes / (tot["seed"] / 1e3) / 1e9 if tot["seed"] > 0 else 0.0
    sa_bytes = 64.0 * cnt[2] + 8.0 * cnt[1]
    # DRAM bytes per launch from an ncu --set full capture of THIS configuration (keyed by reference length and
    # reads per batch); null when no capture of the configuration being run has been committed
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "seed_traffic.json")))
        traffic = tj.get("k_seed", {}).get(f"ref_len={args.ref_len},reads_per_batch={all_batches[0][1]}", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "k_seed", "bound": "hbm", "achieved": seed_gbs, "peak": peak, "unit": "GB/s",
                "frac": seed_gbs / peak, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_read": seed_bytes / n_reads, "launches": len(all_batches),
                "ms_per_launch": tot["seed"] / len(all_batches),
                "sa_kernel": {"bytes_per_read": sa_bytes / n_reads, "GB/s": sa_bytes / (tot["sa"] / 1e3) / 1e9 if tot["sa"] > 0 else 0.0}}
    cells = float(cnt[3] + cnt[5] + cnt[7])
    dp_ms = tot["extend"] + tot["rescue"] + tot["finalize"]
    # SW kernels against the INT32 issue roofline: GCUPS of the kernel proper (its CUDA events) x the reference
    # recurrence's integer operations per cell (ksw_extend2 14, ksw_global2 12, ksw_u8/i16 10) over the
    # MEASURED full-rate integer issue peak (scripts/int_peak.cu on this box: profiles/r2/int_peaks.json, IADD3/IMNMX;
    # DPX add-max / max3 forms issue at half that rate)
    int_peak = None
    try:
        int_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r2", "int_peaks.json")))["iadd3"]) * 1e12
    except Exception:
        pass

    def sw_kernel(cells_k, ms, ops):
        g = cells_k / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"GCUPS": g, "ms_per_step": ms, "ops_per_cell": ops, "frac_of_int_peak": (g * 1e9 * ops / int_peak) if int_peak else None}
    sw = {"cells_per_read": {"extend": cnt[3] / n_reads, "global": cnt[5] / n_reads, "local": cnt[7] / n_reads},
          "GCUPS_stage": cells / (dp_ms / 1e3) / 1e9 if dp_ms > 0 else 0.0,
          "int_peak_lane_ops_per_s": int_peak, "int_peak_source": "measured: profiles/r2/int_peaks.json (scripts/int_peak.cu)",
          "kernels": {"k_xext_run (ksw_extend2, first band try of every chain's best seed)": sw_kernel(cnt[9], tot["k_xext"], 14),
                      "k_resc_sw (ksw_u8/i16 of mem_matesw)": sw_kernel(cnt[10], tot["k_resc_sw"], 10),
                      "k_galn* (ksw_global2 of mem_reg2aln)": sw_kernel(cnt[11], tot["k_galn"], 12)}}

    out = {"metric": metric, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "int32", "data": "synthetic",
           "config": {"workload": workload, "batches_per_step": len(all_batches), "streams_per_gpu": n_workers,
                      "l2": "every step streams the whole job (reads + per-batch stage buffers, > 1 GB) through the GPU, far more than the "
                            "126 MB L2; the 100 MB BWT itself stays largely L2-resident at this reference size (see roofline.traffic)",
                      "parallelism": f"one job, -K batches dealt round-robin over {world} GPU(s); the K passes of the timed region are "
                                     "queued back to back (pass p batch b -> rank (p*B+b) mod N), no drain between passes; e2e drains every call",
                      "value_is": "device-resident reads, SAM text left on the device (no per-batch H2D of reads / D2H of SAM); e2e has both"},
           "gpu_launches": launches, "roofline": roofline, "sw": sw,
           "stage_ms_per_step": {k: v for k, v in tot.items()}, "clocks": sampler.summary()}

    if rank == 0 and not args.no_cpu and os.access(ORACLE, os.X_OK) and world == 1:
        sample = min(args.pairs, args.cpu_sample_pairs or 200_000)
        v, dt, cpu_s = cpu_reference(args, prefix, f1, f2, sample, host_cores)
        out["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": host_cores, "kind": "reference",
                               "sample": f"first {sample} pairs, oracle/_ref/bwa mem -t {host_cores} -K {args.K}, wall {dt:.1f}s"}

    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if world > 1:
        dist.barrier()
    if not args.no_e2e and rank != 0:
        # rank 0's single call drives every GPU of the job; wait for it on the rendezvous store (a NCCL barrier would
        # keep a spinning kernel on this rank's GPU while the product call is being timed)
        import datetime
        store.wait(["b2_e2e_done"], datetime.timedelta(minutes=30))
    if not args.no_e2e and rank == 0:
        # ---- end to end through the reference-facing entry point (what the JNI symbol forwards to) ----
        sam = os.path.join(d, "e2e.sam")
        argv = ["bwa", "mem", "-v", "1", "-K", str(args.K), prefix, f1] + ([f2] if args.paired else [])
        arr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])
        os.environ["B200BWA_DEVICES"] = ",".join(str(g) for g in range(world))
        e2e_pairs = args.pairs
        if lib.b200bwa_main_to(len(argv), arr, sam.encode()) != 0:  # warm-up: loads + caches the index in HBM (all devices)
            raise SystemExit("e2e failed: " + lib.b200bwa_last_error().decode())
        for _ in range(2):                                          # more warm-ups: per-worker buffers at their steady size
            lib.b200bwa_main_to(len(argv), arr, sam.encode())
        lib.b200bwa_process_counters(pc, 1)
        n_e2e = max(1, min(5, args.steps))
        per_call = []
        for _ in range(n_e2e):
            try:  # a fresh output file per call, as SparkBWA gives every task (truncating 780 MB of tmpfs pages costs ~30 ms)
                os.remove(sam)
            except OSError:
                pass
            t0 = time.time()
            if lib.b200bwa_main_to(len(argv), arr, sam.encode()) != 0:
                raise SystemExit("e2e failed: " + lib.b200bwa_last_error().decode())
            per_call.append(time.time() - t0)
        # median of the calls: a single call is a fraction of a second, so one host hiccup (another tenant of the box, a
        # page-cache flush) would otherwise dominate the figure; every call's time is reported next to it
        dt = float(np.median(per_call))
        lib.b200bwa_process_counters(pc, 0)
        out["e2e"] = {"value": n_reads / dt, "unit": "reads/s", "h2d_bytes_per_step": int(pc[0] // n_e2e),
                      "d2h_bytes_per_step": int(pc[1] // n_e2e),
                      "path": f"one b200bwa_main_to call on {world} GPU(s): FASTQ files (tmpfs) -> one SAM file (tmpfs)",
                      "sam_bytes": os.path.getsize(sam), "seconds_per_step": dt,
                      "seconds_per_call": [round(x, 4) for x in per_call], "statistic": "median of the calls"}
        try:
            os.remove(sam)
        except OSError:
            pass
        if world > 1:
            # Companion figure: the job as SparkBWA itself would run it on this box -- N partitions, N concurrent calls
            # of the entry point from N task threads (each leases one GPU: B200BWA_DEVICES_PER_CALL=1), N SAM files.
            # Each partition's SAM equals the reference's on that partition; no single output file to funnel through.
            import threading as _th
            parts = []
            for fsrc in [f1] + ([f2] if args.paired else []):
                with open(fsrc, "rb") as fi:
                    rec = sum(len(fi.readline()) for _ in range(4))
                    per = -(-args.pairs // world)
                    names_p = []
                    for p_ in range(world):
                        fi.seek(rec * p_ * per)
                        fp_ = f"{fsrc}.part{p_}"
                        with open(fp_, "wb") as fo:
                            fo.write(fi.read(rec * max(0, min(per, args.pairs - p_ * per))))
                        names_p.append(fp_)
                    parts.append(names_p)
            os.environ["B200BWA_DEVICES_PER_CALL"] = "1"
            rcs = [0] * world

            def run_part(p_):
                av = ["bwa", "mem", "-v", "1", "-K", str(args.K), prefix] + [x[p_] for x in parts]
                ar = (ctypes.c_char_p * len(av))(*[a.encode() for a in av])
                rcs[p_] = lib.b200bwa_main_to(len(av), ar, f"{sam}.part{p_}".encode())
            tcalls = []
            for it in range(3):
                for p_ in range(world):
                    try:
                        os.remove(f"{sam}.part{p_}")
                    except OSError:
                        pass
                ths = [_th.Thread(target=run_part, args=(p_,)) for p_ in range(world)]
                t0 = time.time()
                [t.start() for t in ths]; [t.join() for t in ths]
                tcalls.append(time.time() - t0)
                if any(rcs):
                    raise SystemExit("e2e (partitions) failed: " + lib.b200bwa_last_error().decode())
            del os.environ["B200BWA_DEVICES_PER_CALL"]
            out["e2e_partitions"] = {"value": n_reads / min(tcalls[1:]), "unit": "reads/s", "seconds_per_round": [round(x, 4) for x in tcalls],
                                     "note": f"the same reads as {world} partitions, {world} concurrent calls (one GPU each), {world} SAM files"}
            for p_ in range(world):
                for fp_ in [f"{sam}.part{p_}"] + [x[p_] for x in parts]:
                    try:
                        os.remove(fp_)
                    except OSError:
                        pass
        if world > 1 and not args.no_e2e_large:
            # Companion figure: the same call on a job N times as long (N x pairs, distinct reads) -- what a
            # configs[2]-sized job sees once start-up and drain are amortised over hundreds of batches.  Not the
            # strong-scaling headline; reported next to it.
            from sparkbwa_b200 import synth
            ctg2 = synth.make_reference(args.ref_len, n_contigs=4 if args.ref_len >= 20_000_000 else 2, seed=args.seed, repeat_frac=0.02)
            g1, g2 = os.path.join(d, "big_1.fq"), os.path.join(d, "big_2.fq")
            big = world * args.pairs
            with open(g1, "wb") as o1, open(g2, "wb") as o2:
                base = 0
                gen = synth.simulate_pairs_fast(ctg2, big, 150, seed=args.seed * 1000 + 23, sub=0.01, indel=0.001) if args.paired else \
                    synth.simulate_pairs_fast(ctg2, big, 250, seed=args.seed * 1000 + 29, sub=0.035, indel=0.015, indel_rounds=6, ins_mean=600.0, ins_sd=60.0)
                for a, b in gen:
                    synth.write_fastq_fixed(o1, [a], base, b"/1" if args.paired else b"")
                    if args.paired:
                        synth.write_fastq_fixed(o2, [b], base, b"/2")
                    base += len(a)
            del ctg2
            argv2 = ["bwa", "mem", "-v", "1", "-K", str(args.K), prefix, g1] + ([g2] if args.paired else [])
            arr2 = (ctypes.c_char_p * len(argv2))(*[a.encode() for a in argv2])
            calls = []
            for _ in range(2):
                try:
                    os.remove(sam)
                except OSError:
                    pass
                t0 = time.time()
                if lib.b200bwa_main_to(len(argv2), arr2, sam.encode()) != 0:
                    raise SystemExit("e2e (large job) failed: " + lib.b200bwa_last_error().decode())
                calls.append(time.time() - t0)
            nr = big * (2 if args.paired else 1)
            out["e2e_large_job"] = {"value": nr / min(calls), "unit": "reads/s", "reads": nr, "seconds_per_call": [round(x, 4) for x in calls],
                                    "sam_bytes": os.path.getsize(sam), "note": f"{world} x the job of `e2e`, one call on {world} GPUs"}
            for p in (g1, g2, sam):
                try:
                    os.remove(p)
                except OSError:
                    pass
        if store is not None:
            store.set("b2_e2e_done", "1")
    if rank == 0:
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def prepare_reference_cpu_only(args, d):
    """--impl reference on a box where the engine arm has not run yet: the index is built by the
    reference's own `bwa index` (this arm is the reference; it may use its own tools)."""
    from sparkbwa_b200 import synth
    prefix = os.path.join(d, "ref")
    n_ctg = 4 if args.ref_len >= 20_000_000 else 2
    ctg = synth.make_reference(args.ref_len, n_contigs=n_ctg, seed=args.seed, repeat_frac=0.02)
    if not os.path.exists(prefix + ".sa"):
        fa = os.path.join(d, "ref.fa")
        synth.write_fasta(fa, ctg)
        t = time.time()
        subprocess.run([ORACLE, "index", "-p", prefix + ".tmp", fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for e in ("pac", "ann", "amb", "bwt", "sa"):
            os.replace(prefix + ".tmp." + e, prefix + "." + e)
        sys.stderr.write(f"[bench] reference index built by bwa index in {time.time() - t:.1f}s\n")
    return ctg, prefix


if __name__ == "__main__":
    main()
