#!/usr/bin/env python
"""Differential fuzzing of the engine against the oracle (test infrastructure, CPU only).

Every trial draws a reference, a read set and a `bwa mem` option string from a seeded generator, runs the unmodified
reference (`oracle/_ref/bwa`) and the host build of the kernel bodies (`tests/hostsim/_build/b200bwa_hostsim`, the same
.cuh code as libbwa.so compiled for the CPU) on the same files and compares the SAM bodies line by line.  A mismatch
keeps its directory and is appended to the report; the findings of a campaign go into tests/ as fixed cases.

  python scripts/fuzz_parity.py --first 1000 --count 200 --jobs 6 --out /tmp/b2fuzz
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402
from sparkbwa_b200 import synth  # noqa: E402


def draw_extreme(seed: int):
    """Corner-heavy trials: tiny -K batches (insert-size statistics fail or flip between batches), degenerate scoring,
    very short seeds, N-rich reads, tiny and highly repetitive references, long reads under every preset."""
    r = np.random.default_rng(seed)
    pick = lambda xs: xs[int(r.integers(0, len(xs)))]  # noqa: E731
    rep = float(pick([0.0, 0.3, 0.8, 0.95]))
    ref_kw = dict(total_len=int(pick([3_000, 20_000, 80_000, 200_000])), n_contigs=int(pick([1, 2, 7])), seed=seed * 7 + 1,
                  repeat_frac=rep)
    if rep > 0:
        ref_kw.update(rep_len=(int(pick([20, 60, 200])), int(pick([250, 800, 2500]))), rep_copies=(2, int(pick([10, 80, 300]))),
                      rep_div=float(pick([0.0, 0.0, 0.02, 0.1])))
    if ref_kw["total_len"] >= 20_000 and r.random() < 0.25:
        ref_kw["alt"] = dict(n_alt=int(pick([1, 4, 15])), alt_len=(300, 2500), div=float(pick([0.0, 0.01, 0.06])), seed=seed * 7 + 2)
    max_len = max(20, ref_kw["total_len"] // ref_kw["n_contigs"] // 4)
    read_len = min(max_len, int(pick([20, 25, 33, 70, 101, 150, 151, 251, 400, 1000, 2500, 4000])))
    paired = bool(r.random() < 0.6) and read_len <= 400
    n = int(pick([1, 7, 60, 300])) if read_len <= 400 else int(pick([3, 20, 40]))
    reads_kw = dict(n=n, read_len=read_len, paired=paired, seed=seed * 7 + 3, sub=float(pick([0.0, 0.01, 0.04, 0.12, 0.2])),
                    indel=float(pick([0.0, 0.002, 0.02, 0.06])), n_frac=float(pick([0.0, 0.01, 0.1, 0.4])),
                    junk_frac=float(pick([0.0, 0.1, 0.5])), chimeric_frac=float(pick([0.0, 0.2, 0.6])) if read_len >= 70 else 0.0,
                    ragged=bool(r.random() < 0.3))
    if paired:
        reads_kw.update(ins_mean=float(pick([read_len + 11, 2 * read_len, 500, 1500])), ins_sd=float(pick([1, 30, 300])))
        if r.random() < 0.4:
            reads_kw["mate2_sub"] = float(pick([0.0, 0.08, 0.2]))
    opts = []
    pool = [
        (0.3, lambda: ["-a"]), (0.2, lambda: ["-M"]), (0.2, lambda: ["-Y"]), (0.15, lambda: ["-S"]), (0.15, lambda: ["-P"]),
        (0.15, lambda: ["-j"]), (0.1, lambda: ["-C"]), (0.1, lambda: ["-V"]), (0.1, lambda: ["-1"]),
        (0.5, lambda: ["-K", str(int(pick([300, 2_000, 9_000, 40_000])))]),
        (0.3, lambda: ["-k", str(int(pick([5, 8, 12, 19, 40, 100])))]),
        (0.3, lambda: ["-w", str(int(pick([0, 1, 3, 50, 500])))]),
        (0.2, lambda: ["-d", str(int(pick([0, 1, 10, 1000])))]),
        (0.2, lambda: ["-r", pick(["0.5", "1.0", "1.01", "4", "100"])]),
        (0.2, lambda: ["-y", str(int(pick([0, 1, 20, 100000])))]),
        (0.2, lambda: ["-c", str(int(pick([1, 2, 10, 50, 100000])))]),
        (0.15, lambda: ["-D", pick(["0.0", "0.5", "0.99", "1.0"])]),
        (0.15, lambda: ["-W", str(int(pick([0, 5, 60, 300])))]),
        (0.15, lambda: ["-m", str(int(pick([0, 1, 2, 1000])))]),
        (0.3, lambda: ["-A", str(int(pick([1, 2, 5, 10])))]),
        (0.3, lambda: ["-B", str(int(pick([0, 1, 3, 8, 20])))]),
        (0.3, lambda: ["-O", pick(["0", "0,0", "1,9", "12", "30,2"])]),
        (0.3, lambda: ["-E", pick(["1", "4", "1,5", "7,1"])]),
        (0.3, lambda: ["-L", pick(["0", "0,20", "25,0", "100", "1,1"])]),
        (0.2, lambda: ["-U", str(int(pick([0, 1, 30, 200])))]),
        (0.3, lambda: ["-T", str(int(pick([0, 1, 15, 60, 200])))]),
        (0.15, lambda: ["-h", pick(["0", "1,1", "50", "200,400"])]),
        (0.1, lambda: ["-Q", str(int(pick([0, 60, 255])))]),
        (0.15, lambda: ["-G", str(int(pick([0, 10, 100, 10000000])))]),
        (0.15, lambda: ["-N", str(int(pick([0, 1, 100])))]),
        (0.15, lambda: ["-s", str(int(pick([0, 1, 300])))]),
        (0.15, lambda: ["-X", pick(["0.0", "0.5", "1.0"])]),
        (0.1, lambda: ["-R", r"@RG\tID:fz\tSM:s\tPL:x"]),
        (0.1, lambda: ["-H", "@CO\tfuzz"]),
        (0.2, lambda: ["-I", pick(["100", "400,1", "350,600", "300,30,301", "2000,500,9000,1"])]),
        (0.25, lambda: ["-x", pick(["pacbio", "ont2d", "intractg"])]),
        (0.1, lambda: ["-t", str(int(pick([1, 3])))]),
    ]
    for p, mk in pool:
        if r.random() < p:
            opts += mk()
    inter = paired and r.random() < 0.2
    return ref_kw, reads_kw, opts, inter


def draw(seed: int):
    """(reference kwargs, reads kwargs, option list, interleaved?) of trial `seed` (seeds >= 1 000 000: draw_extreme)."""
    if 1_000_000 <= seed < 3_000_000:
        return draw_extreme(seed)
    r = np.random.default_rng(seed)
    pick = lambda xs: xs[int(r.integers(0, len(xs)))]  # noqa: E731
    rep = float(pick([0.0, 0.02, 0.05, 0.2, 0.5, 0.7]))
    ref_kw = dict(total_len=int(pick([60_000, 150_000, 300_000])), n_contigs=int(pick([1, 2, 3, 5])), seed=seed * 7 + 1,
                  repeat_frac=rep)
    if rep >= 0.2:
        ref_kw.update(rep_len=(int(pick([50, 100, 300])), int(pick([400, 1500, 3000]))), rep_copies=(2, int(pick([6, 20, 60]))),
                      rep_div=float(pick([0.0, 0.01, 0.04, 0.08])))
    if r.random() < 0.2:
        ref_kw["alt"] = dict(n_alt=int(pick([2, 6, 12])), alt_len=(800, 3000), div=float(pick([0.005, 0.02, 0.05])), seed=seed * 7 + 2)
    read_len = int(pick([30, 36, 50, 76, 100, 125, 150, 150, 200, 250, 300, 500, 800]))
    paired = bool(r.random() < 0.6)
    n = int(pick([150, 300, 600])) if read_len <= 250 else int(pick([40, 80, 120]))
    reads_kw = dict(n=n, read_len=read_len, paired=paired, seed=seed * 7 + 3, sub=float(pick([0.0, 0.005, 0.01, 0.02, 0.05, 0.09])),
                    indel=float(pick([0.0, 0.001, 0.004, 0.015, 0.03])), n_frac=float(pick([0.0, 0.0, 0.002, 0.02])),
                    junk_frac=float(pick([0.0, 0.01, 0.05])), chimeric_frac=float(pick([0.0, 0.0, 0.1, 0.3])) if read_len >= 70 else 0.0,
                    ragged=bool(r.random() < 0.15))
    if paired:
        reads_kw.update(ins_mean=float(pick([read_len + 20, 250, 400, 700])), ins_sd=float(pick([5, 40, 120])))
        if r.random() < 0.4:
            reads_kw["mate2_sub"] = float(pick([0.05, 0.1, 0.15]))
    opts = []
    pool = [
        (0.25, lambda: ["-a"]), (0.15, lambda: ["-M"]), (0.15, lambda: ["-Y"]), (0.1, lambda: ["-S"]), (0.1, lambda: ["-P"]),
        (0.1, lambda: ["-j"]), (0.1, lambda: ["-C"]), (0.1, lambda: ["-V"]), (0.1, lambda: ["-1"]),
        (0.25, lambda: ["-K", str(int(pick([20_000, 50_000, 200_000])))]),
        (0.15, lambda: ["-k", str(int(pick([10, 15, 17, 23, 30])))]),
        (0.15, lambda: ["-w", str(int(pick([5, 20, 40, 200])))]),
        (0.15, lambda: ["-d", str(int(pick([20, 50, 200])))]),
        (0.15, lambda: ["-r", pick(["1.0", "1.2", "2.5", "10"])]),
        (0.1, lambda: ["-y", str(int(pick([0, 4, 8, 50])))]),
        (0.15, lambda: ["-c", str(int(pick([5, 20, 100, 2000])))]),
        (0.1, lambda: ["-D", pick(["0.1", "0.3", "0.8"])]),
        (0.1, lambda: ["-W", str(int(pick([10, 30, 40])))]),
        (0.1, lambda: ["-m", str(int(pick([0, 3, 10, 100])))]),
        (0.15, lambda: ["-A", str(int(pick([1, 2, 3])))]),
        (0.15, lambda: ["-B", str(int(pick([1, 2, 4, 6, 9])))]),
        (0.15, lambda: ["-O", pick(["0", "1", "4", "6,8", "10,3"])]),
        (0.15, lambda: ["-E", pick(["1", "2", "1,3", "3,1"])]),
        (0.15, lambda: ["-L", pick(["0", "3", "5,9", "12,0", "40"])]),
        (0.15, lambda: ["-U", str(int(pick([0, 5, 9, 17, 40])))]),
        (0.2, lambda: ["-T", str(int(pick([0, 10, 20, 30, 45])))]),
        (0.1, lambda: ["-h", pick(["1", "3", "5,300", "20"])]),
        (0.1, lambda: ["-Q", str(int(pick([0, 10, 30])))]),
        (0.1, lambda: ["-G", str(int(pick([50, 1000, 100000])))]),
        (0.1, lambda: ["-N", str(int(pick([1, 5, 25])))]),
        (0.1, lambda: ["-s", str(int(pick([1, 5, 30])))]),
        (0.1, lambda: ["-X", pick(["0.3", "0.7", "0.95"])]),
        (0.1, lambda: ["-R", r"@RG\tID:fz\tSM:s\tPL:x"]),
        (0.1, lambda: ["-I", pick(["300", "400,40", "350,60,900", "300,30,700,100"])]),
        (0.08, lambda: ["-x", pick(["pacbio", "ont2d", "intractg"])]),
    ]
    for p, mk in pool:
        if r.random() < p:
            opts += mk()
    inter = paired and r.random() < 0.15
    return ref_kw, reads_kw, opts, inter


def write_mangled(path: str, reads, r, style):
    """FASTQ / FASTA text the way real files vary (kseq semantics, bwa/kseq.h:175-215): CR/LF, blank lines, '+name' lines,
    comments after the name, lower case and ambiguity codes, empty records, FASTA records, a missing final newline, gzip.
    (No '-' in reads: nst_nt4_table gives it code 5, with which the reference indexes past a row of its 5x5 scoring matrix
    and prints a NUL byte into SEQ -- undefined input, not reproduced.)"""
    import gzip
    eol = b"\r\n" if style["crlf"] else b"\n"
    out = bytearray()
    for k, (name, seq, qual) in enumerate(reads):
        if style["comment"] and r.random() < 0.5:
            name = name + (b" " if r.random() < 0.7 else b"\t") + b"BC:Z:ACGT%d" % k
        if style["lower"] and r.random() < 0.3:
            seq = seq.lower()
        if style["iupac"] and len(seq) > 3 and r.random() < 0.2:
            b = bytearray(seq); b[int(r.integers(0, len(b)))] = ord(str(r.choice(list("RYKMSW.nX*")))); seq = bytes(b)
        if style["empty"] and r.random() < 0.03:
            seq, qual = b"", b""
        if style.get("badqual") and len(qual) > 8 and r.random() < 0.03:   # kseq_read returns -2: the batch ends, reading resumes behind it
            qual = qual + b"IIIII" if r.random() < 0.5 else qual[:-3]
        if style["fasta"] == 2 or (style["fasta"] == 1 and k % 3 == 1):
            out += b">" + name + eol
            w = style["fold"]
            for p in range(0, max(len(seq), 1), w):
                out += seq[p:p + w] + eol
        else:
            out += b"@" + name + eol + seq + eol + (b"+" + (name.split()[0] if style["plusname"] else b"")) + eol + qual + eol
        if style["blank"] and r.random() < 0.1:
            out += eol
    if style["no_final_eol"] and out.endswith(eol):
        del out[-len(eol):]
    if style.get("truncate") and len(out) > 400:   # a file cut off in mid-record (kseq returns -2 / EOF: the reader stops there)
        del out[len(out) - int(r.integers(1, 400)):]
    if style["gz"]:
        path += ".gz"
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(bytes(out))
    else:
        with open(path, "wb") as f:
            f.write(bytes(out))
    return path


def draw_style(r):
    pr = lambda p: bool(r.random() < p)  # noqa: E731
    return dict(crlf=pr(0.2), comment=pr(0.3), lower=pr(0.2), iupac=pr(0.2), empty=pr(0.15), fasta=int(r.choice([0, 0, 0, 1, 2])),
                fold=int(r.choice([60, 7, 100000])), plusname=pr(0.2), blank=pr(0.3), no_final_eol=pr(0.25), gz=pr(0.15), truncate=pr(0.2), badqual=pr(0.2))


def trial(args):
    seed, outdir, keep, engine = args
    ref_kw, reads_kw, opts, inter = draw(seed)
    d = os.path.join(outdir, f"t{seed}")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    rec = dict(seed=seed, ref=ref_kw, reads=reads_kw, opts=opts, interleaved=inter)
    t0 = time.time()
    try:
        kw = dict(ref_kw)
        alt_kw = kw.pop("alt", None)
        ctg = synth.make_reference(**kw)
        alt = []
        if alt_kw:
            ctg, alt = synth.add_alt_contigs(ctg, **alt_kw)
        fa = os.path.join(d, "ref.fa")
        synth.write_fasta(fa, ctg)
        if alt:
            with open(fa + ".alt", "w") as f:
                f.write("@SQ\tSN:chr1\tLN:1\n")
                for nm in alt:
                    f.write(f"{nm}\t0\tchr1\t1\t60\t100M\t*\t0\t0\t*\t*\n")
        if 1_000_000 <= seed < 3_000_000 and seed % 4 == 0:   # ambiguity codes / lower case in the REFERENCE (.amb holes, random fill)
            mr = np.random.default_rng(seed + 5)
            lines = open(fa, "rb").read().split(b"\n")
            for k, ln in enumerate(lines):
                if ln and not ln.startswith(b">") and mr.random() < 0.02:
                    b = bytearray(ln)
                    a0 = int(mr.integers(0, len(b))); m = int(mr.choice([1, 3, 60]))
                    b[a0:a0 + m] = bytes([ord(str(mr.choice(list("NNNRYn"))))]) * len(b[a0:a0 + m])
                    lines[k] = bytes(b)
                elif ln and not ln.startswith(b">") and mr.random() < 0.05:
                    lines[k] = ln.lower()
            open(fa, "wb").write(b"\n".join(lines))
            rec["ref_mangled"] = True
        pu.run([pu.ORACLE, "index", fa])
        try:
            r1, r2 = synth.simulate_reads(ctg, **reads_kw)
        except ValueError as e:  # the draw asked for reads longer than a contig: not a trial
            rec.update(ok=True, skipped=repr(e)[:100], seconds=0)
            shutil.rmtree(d, ignore_errors=True)
            return rec
        fq = [os.path.join(d, "r_1.fq")]
        if 3_000_000 <= seed < 4_000_000:   # "format" trials: the content of a regular trial in irregular files
            fr = np.random.default_rng(seed + 17)
            style = draw_style(fr)
            rec["style"] = [dict(style)]
            fq[0] = write_mangled(fq[0], r1, fr, style)
            if r2 is not None:
                if fr.random() < 0.5:   # the mates' files need not be written alike
                    style = draw_style(fr)
                    style["empty"] = False
                rec["style"].append(dict(style))
                if fr.random() < 0.1:
                    r2 = r2[:max(1, len(r2) - int(fr.integers(1, 4)))]   # the 2nd file has fewer sequences (bwa/bwa.c:66-69)
                fq.append(write_mangled(os.path.join(d, "r_2.fq"), r2, fr, style))
        else:
            synth.write_fastq(fq[0], r1)
            if r2 is not None:
                fq.append(os.path.join(d, "r_2.fq"))
                synth.write_fastq(fq[1], r2)
        case = dict(dir=d, ref=fa, opts=opts, paired=reads_kw["paired"], fq=fq)
        if inter and not (3_000_000 <= seed < 4_000_000):
            case["fq"] = [pu.make_interleaved(case)]
            case["opts"] = opts + ["-p"]
        ro = subprocess.run([pu.ORACLE, "mem"] + case["opts"] + [fa] + case["fq"], stdout=open(os.path.join(d, "oracle.sam"), "wb"),
                            stderr=subprocess.PIPE, timeout=1200)
        if engine == "gpu":  # the product binary on cuda:0 (two stream sets: many short-lived processes share the GPU)
            binary, env = pu.CLI, dict(os.environ, B200BWA_WORKERS=os.environ.get("B200BWA_WORKERS", "2"))
        else:
            binary, env = pu.HOSTSIM, dict(os.environ, **pu.HOSTSIM_ENV)
        for kv in filter(None, os.environ.get("FUZZ_ENGINE_ENV", "").split(",")):   # e.g. tiny capacities: every overflow / retry path
            env[kv.split("=", 1)[0]] = kv.split("=", 1)[1]
        re_ = subprocess.run([binary, "mem"] + case["opts"] + ["-f", os.path.join(d, "engine.sam"), fa] + case["fq"],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=120 if engine == "gpu" else 2400)
        rec["rc_oracle"], rec["rc_engine"] = ro.returncode, re_.returncode
        if ro.returncode < 0:  # the reference itself died of a signal (it does with some degenerate scorings): nothing to compare
            rec.update(ok=True, skipped="oracle killed by signal %d" % -ro.returncode)
        elif ro.returncode != 0 or re_.returncode != 0:
            # both refusing the same option string is agreement; one side alone is a finding
            rec["ok"] = (ro.returncode != 0) == (re_.returncode != 0)
            rec["stderr_engine"] = re_.stderr.decode(errors="replace")[-600:]
            rec["stderr_oracle"] = ro.stderr.decode(errors="replace")[-300:]
        else:
            diff = pu.first_diff(os.path.join(d, "oracle.sam"), os.path.join(d, "engine.sam"))
            rec["ok"] = not diff
            rec["lines"] = len(pu.sam_body(os.path.join(d, "oracle.sam")))
            if diff:
                rec["diff"] = diff
    except Exception as e:  # noqa: BLE001
        rec["ok"] = False
        rec["error"] = repr(e)[:600]
    rec["seconds"] = round(time.time() - t0, 1)
    if rec["ok"] and not keep:
        shutil.rmtree(d, ignore_errors=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=1000)
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--out", default="/tmp/b2fuzz")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--engine", choices=["hostsim", "gpu"], default="hostsim",
                    help="hostsim: kernel bodies compiled for the CPU; gpu: sparkbwa_b200/b200bwa (libbwa.so) on cuda:0")
    ap.add_argument("--seconds", type=float, default=0, help="stop handing out trials after this many seconds (0: run --count)")
    a = ap.parse_args()
    assert pu.have_oracle() and os.access(pu.CLI if a.engine == "gpu" else pu.HOSTSIM, os.X_OK), "build oracle/_ref and the engine first"
    os.makedirs(a.out, exist_ok=True)
    rep = os.path.join(a.out, "report.jsonl")
    n_ok = n_bad = n_lines = 0
    t0 = time.time()
    with ProcessPoolExecutor(a.jobs) as ex, open(rep, "a") as f:
        nxt, end = a.first, a.first + a.count
        while nxt < end and not (a.seconds and time.time() - t0 > a.seconds):
            chunk = range(nxt, min(end, nxt + 4 * a.jobs))
            nxt = chunk[-1] + 1
            for rec in ex.map(trial, [(s, a.out, a.keep, a.engine) for s in chunk]):
                f.write(json.dumps(rec) + "\n")
                f.flush()
                n_ok += rec["ok"]
                n_bad += not rec["ok"]
                n_lines += rec.get("lines", 0)
                if not rec["ok"]:
                    print("MISMATCH", json.dumps(rec)[:1500], flush=True)
    print(f"engine {a.engine}: trials {n_ok + n_bad} (seeds {a.first}..{nxt - 1}): {n_ok} identical, {n_bad} findings, "
          f"{n_lines} SAM lines compared in {time.time() - t0:.0f} s; report {rep}")


if __name__ == "__main__":
    main()
