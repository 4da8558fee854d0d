#!/usr/bin/env python
"""BASELINE.json configs[2]-class run: paired-end 2x150 bp reads against a >= 1 Gbp synthetic reference, enough pairs
for the pair index to cross 2^23 (the int32 wrap of `id<<8` in mem_pair), one job over all visible GPUs.

  python scripts/cfg2_run.py --ref-len 1000000000 --pairs 8600000 --out gpurun_out/r2e

Steps (each timed, all reported in <out>/cfg2_L<ref-len>_P<pairs>.json):
  1. synthetic reference (24 contigs at 3.1 Gbp, 8 below) -> FASTA -> `b200bwa index` (GPU suffix sort)
  2. reads (sparkbwa_b200.synth.simulate_pairs_fast) -> two FASTQ files
  3. one b200bwa_main_to call (cold: index load + fan-out; then warm) -> SAM
  4. parity: the reference's `bwa mem -t <cores>` on the first --prefix-batches batches (same -K, so the same batch
     cuts and insert-size statistics) against the head of the SAM; and the reference's mem_process_seqs
     (oracle/_ref/oracle_batch) on late batches with their absolute n_processed against the matching SAM records.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparkbwa_b200 import synth, index  # noqa: E402

ORACLE = os.path.join(ROOT, "oracle", "_ref", "bwa")
ORACLE_BATCH = os.path.join(ROOT, "oracle", "_ref", "oracle_batch")
LIB = os.path.join(ROOT, "sparkbwa_b200", "libbwa.so")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=1_000_000_000)
    ap.add_argument("--pairs", type=int, default=8_600_000)
    ap.add_argument("--K", type=int, default=10_000_000)
    ap.add_argument("--prefix-batches", type=int, default=12)
    ap.add_argument("--late-batches", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/cfg2")
    ap.add_argument("--work", default="/dev/shm/b2cfg2")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--sim-procs", type=int, default=0, help="processes simulating reads (0: half the cores, at most 16)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    os.makedirs(a.work, exist_ok=True)
    rep = {"ref_len": a.ref_len, "pairs": a.pairs, "K": a.K, "host_cores": os.cpu_count()}
    st = os.statvfs(a.work)
    rep["work_free_gb"] = st.f_bavail * st.f_frsize / 1e9

    def log(msg):
        sys.stderr.write(f"[cfg2 +{time.time() - t_start:7.1f}s] {msg}\n")
        sys.stderr.flush()
    t_start = time.time()

    # ---- 1. reference + index ------------------------------------------------------------------
    n_ctg = 24 if a.ref_len >= 2_000_000_000 else 8
    t = time.time()
    ctg = synth.make_reference(a.ref_len, n_contigs=n_ctg, seed=1, repeat_frac=0.02)
    rep["make_reference_s"] = time.time() - t
    log(f"reference of {a.ref_len} bp in {n_ctg} contigs generated")
    prefix = os.path.join(a.work, "ref")
    t = time.time()
    index.build_index(ctg, prefix)
    rep["index_build_s"] = time.time() - t      # FASTA write + parse + pack + GPU BWT/SA + file writes
    rep["index_bytes"] = {e: os.path.getsize(prefix + "." + e) for e in ("bwt", "sa", "pac")}
    log(f"index built in {rep['index_build_s']:.1f}s: {rep['index_bytes']}")

    # ---- 2. reads --------------------------------------------------------------------------------
    f1, f2 = os.path.join(a.work, "r_1.fq"), os.path.join(a.work, "r_2.fq")
    t = time.time()
    # simulated in slices by forked workers (each slice its own seed), concatenated in order
    n_proc = max(1, min(a.sim_procs or (os.cpu_count() or 1) // 2, 16))
    per = -(-a.pairs // n_proc)
    pids = []
    for k in range(n_proc):
        lo, hi = k * per, min(a.pairs, (k + 1) * per)
        if lo >= hi:
            break
        pid = os.fork()
        if pid == 0:
            with open(f"{f1}.{k}", "wb") as o1, open(f"{f2}.{k}", "wb") as o2:
                base = lo
                for s1, s2 in synth.simulate_pairs_fast(ctg, hi - lo, 150, seed=2024 + k, sub=0.01, indel=0.001):
                    synth.write_fastq_fixed(o1, [s1], base, b"/1")
                    synth.write_fastq_fixed(o2, [s2], base, b"/2")
                    base += len(s1)
            os._exit(0)
        pids.append(pid)
    for pid in pids:
        if os.waitpid(pid, 0)[1] != 0:
            raise SystemExit("read simulation worker failed")
    for f in (f1, f2):
        with open(f, "wb") as fo:
            for k in range(len(pids)):
                with open(f"{f}.{k}", "rb") as fi:
                    while True:
                        buf = fi.read(1 << 26)
                        if not buf:
                            break
                        fo.write(buf)
                os.remove(f"{f}.{k}")
    del ctg
    rep["simulate_reads_s"] = time.time() - t
    rep["fastq_bytes"] = os.path.getsize(f1) + os.path.getsize(f2)
    log(f"{a.pairs} pairs simulated and written in {rep['simulate_reads_s']:.1f}s")

    # ---- 3. the job --------------------------------------------------------------------------------
    lib = ctypes.CDLL(LIB)
    lib.b200bwa_main_to.restype = ctypes.c_int
    lib.b200bwa_last_error.restype = ctypes.c_char_p
    sam = os.path.join(a.work, "ours.sam")
    argv = ["bwa", "mem", "-v", "1", "-K", str(a.K), prefix, f1, f2]
    arr = (ctypes.c_char_p * len(argv))(*[x.encode() for x in argv])
    calls = []
    for k in range(2):
        if os.path.exists(sam):
            os.remove(sam)
        t = time.time()
        rc = lib.b200bwa_main_to(len(argv), arr, sam.encode())
        dt = time.time() - t
        if rc != 0:
            raise SystemExit("engine failed: " + lib.b200bwa_last_error().decode())
        calls.append(dt)
        log(f"engine call {k}: {dt:.2f}s = {2 * a.pairs / dt / 1e6:.2f} M reads/s")
    rep["engine_seconds"] = {"cold_incl_index_load": calls[0], "warm": calls[1]}
    rep["engine_reads_per_s_warm"] = 2 * a.pairs / calls[1]
    rep["sam_bytes"] = os.path.getsize(sam)
    try:
        import torch
        rep["n_gpus"] = torch.cuda.device_count()
    except Exception:
        pass

    # ---- 4. parity -------------------------------------------------------------------------------
    pairs_per_batch = -(-a.K // 300)          # 150 + 150 bases per pair; a batch ends once its bases reach K
    rec_bytes = None
    with open(f1, "rb") as f:
        rec_bytes = sum(len(f.readline()) for _ in range(4))
    n_batches = -(-a.pairs // pairs_per_batch)
    rep["n_batches"] = n_batches
    nb = min(a.prefix_batches, n_batches)
    if nb > 0 and os.access(ORACLE, os.X_OK):
        n_pre = min(a.pairs, nb * pairs_per_batch)
        h1, h2 = os.path.join(a.work, "pre_1.fq"), os.path.join(a.work, "pre_2.fq")
        for src, dst in ((f1, h1), (f2, h2)):
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                fo.write(fi.read(rec_bytes * n_pre))
        t = time.time()
        ref_sam = os.path.join(a.work, "oracle_pre.sam")
        with open(ref_sam, "wb") as fo:
            subprocess.run([ORACLE, "mem", "-v", "1", "-t", str(os.cpu_count() or 1), "-K", str(a.K), prefix, h1, h2], check=True,
                           stdout=fo, stderr=subprocess.DEVNULL)
        dt = time.time() - t
        rep["oracle_prefix"] = {"pairs": n_pre, "batches": nb, "seconds": dt, "reads_per_s": 2 * n_pre / dt, "threads": os.cpu_count()}
        log(f"oracle on the first {nb} batches: {dt:.1f}s = {2 * n_pre / dt / 1e6:.3f} M reads/s")
        same, n_cmp = True, 0
        with open(ref_sam, "rb") as fr, open(sam, "rb") as fo_:
            for lr in fr:
                lo = fo_.readline()
                if lr.startswith(b"@PG"):
                    continue
                n_cmp += 1
                if lr != lo:
                    same = False
                    rep["prefix_first_diff"] = [lr.decode(errors="replace")[:300], lo.decode(errors="replace")[:300]]
                    break
        rep["prefix_parity"] = {"identical": same, "lines_compared": n_cmp}
        log(f"prefix parity: identical={same} over {n_cmp} lines")
        for p in (h1, h2, ref_sam):
            os.remove(p)
    # late batches: the reference's mem_process_seqs with the batch's absolute n_processed
    late = []
    if a.late_batches > 0 and os.access(ORACLE_BATCH, os.X_OK):
        # byte offsets of the SAM records of a late batch: every pair's records start with its read name r%09d
        want = [b for b in range(n_batches - a.late_batches, n_batches) if b >= 0]
        starts = {b: b * pairs_per_batch for b in want}
        ends = {b: min(a.pairs, (b + 1) * pairs_per_batch) for b in want}
        ours = {b: [] for b in want}
        lo_pair, hi_pair = min(starts.values()), max(ends.values())
        with open(sam, "rb") as f:
            # the tail of the file holds the late batches: scan back far enough (records are < 1 KB per read here)
            back = min(os.path.getsize(sam), (hi_pair - lo_pair + pairs_per_batch) * 2 * 1024)
            f.seek(os.path.getsize(sam) - back)
            f.readline()
            for line in f:
                if line.startswith(b"@"):
                    continue
                pid = int(line[1:10])
                for b in want:
                    if starts[b] <= pid < ends[b]:
                        ours[b].append(line)
        for b in want:
            b1, b2 = os.path.join(a.work, "late_1.fq"), os.path.join(a.work, "late_2.fq")
            for src, dst in ((f1, b1), (f2, b2)):
                with open(src, "rb") as fi, open(dst, "wb") as fo:
                    fi.seek(rec_bytes * starts[b])
                    fo.write(fi.read(rec_bytes * (ends[b] - starts[b])))
            t = time.time()
            r = subprocess.run([ORACLE_BATCH, prefix, str(2 * starts[b]), b1, b2], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
            ref_lines = r.stdout.splitlines(True)
            ok = ref_lines == ours[b]
            late.append({"batch": b, "first_pair": starts[b], "pairs": ends[b] - starts[b], "pair_index_ge_2^23": starts[b] >= (1 << 23),
                         "identical": ok, "lines": len(ref_lines), "oracle_seconds": time.time() - t})
            log(f"late batch {b} (first pair {starts[b]}): identical={ok} over {len(ref_lines)} lines")
            for p in (b1, b2):
                os.remove(p)
    rep["late_batches"] = late
    out = os.path.join(a.out, f"cfg2_L{a.ref_len}_P{a.pairs}.json")
    with open(out, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))
    if not a.keep:
        for p in (sam, f1, f2):
            os.remove(p)


if __name__ == "__main__":
    main()
