#!/usr/bin/env python
"""Differential fuzzing of the index builder against `bwa index` (test infrastructure, CPU: the host checker build of
b2_index.cu inside tests/hostsim; `--engine gpu`: sparkbwa_b200/b200bwa on cuda:0).  Random FASTA files with the
features the packer has to reproduce (bwa/bntseq.c:228-328): ambiguity codes and runs of N (replaced by lrand48() & 3
after srand48(11)), lower case, comments after the name, blank and ragged lines, many short contigs, exact repeats.
The five files .pac .ann .amb .bwt .sa must be byte-identical.

  python scripts/fuzz_index.py --first 0 --count 300 --jobs 6
"""
import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402


def make_fasta(path, seed):
    r = np.random.default_rng(seed)
    n_ctg = int(r.choice([1, 1, 2, 5, 40]))
    total = int(r.choice([1, 7, 33, 500, 5_000, 60_000, 400_000]))
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for c in range(n_ctg):
            ln = max(1, total // n_ctg + int(r.integers(-total // (2 * n_ctg) - 1, total // (2 * n_ctg) + 2)))
            s = alphabet[r.integers(0, 4, size=ln)].copy()
            if r.random() < 0.5 and ln > 50:   # exact repeats (deep tie refinement)
                for _ in range(int(r.integers(1, 6))):
                    a, b = sorted(int(x) for x in r.integers(0, ln, size=2))
                    seg = s[a:min(b, a + 3000)]
                    if len(seg) == 0:
                        continue
                    for _ in range(int(r.integers(1, 8))):
                        p = int(r.integers(0, ln))
                        m = min(len(seg), ln - p)
                        s[p:p + m] = seg[:m]
            if r.random() < 0.3 and ln > 10:   # a homopolymer / short-period stretch
                p = int(r.integers(0, ln)); m = int(r.integers(1, min(ln - p, 5000) + 1)); per = int(r.choice([1, 2, 3, 7]))
                s[p:p + m] = np.resize(s[p:p + per], m)
            if r.random() < 0.6:               # ambiguity codes, singly and in runs
                for _ in range(int(r.integers(1, 10))):
                    p = int(r.integers(0, ln)); m = int(r.choice([1, 1, 2, 10, 300]))
                    s[p:p + m] = ord(str(r.choice(list("NNNRYKMSWnx-"))))
            if r.random() < 0.4:
                m = r.random(ln) < 0.3
                s[m] = np.frombuffer(bytes(s[m]).lower(), dtype=np.uint8)
            name = f"ctg{c}".encode() + (b" some comment %d" % c if r.random() < 0.5 else b"") + (b"\tx" if r.random() < 0.2 else b"")
            f.write(b">" + name + b"\n")
            w = int(r.choice([60, 70, 1, 1000000]))
            i = 0
            while i < ln:
                k = w if r.random() > 0.1 else int(r.integers(1, w + 1))
                f.write(bytes(s[i:i + k]) + b"\n")
                i += k
                if r.random() < 0.02:
                    f.write(b"\n")


def trial(args):
    seed, outdir, engine = args
    d = os.path.join(outdir, f"i{seed}")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    fa = os.path.join(d, "ref.fa")
    make_fasta(fa, seed)
    ro = subprocess.run([pu.ORACLE, "index", "-p", os.path.join(d, "o"), fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    binary = pu.CLI if engine == "gpu" else pu.HOSTSIM
    re_ = subprocess.run([binary, "index", "-p", os.path.join(d, "e"), fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if ro.returncode < 0:
        shutil.rmtree(d, ignore_errors=True)
        return seed, True, "oracle killed by signal"
    if ro.returncode != 0 or re_.returncode != 0:
        ok = (ro.returncode != 0) == (re_.returncode != 0)
        msg = f"rc oracle {ro.returncode} engine {re_.returncode}: {re_.stderr.decode(errors='replace')[-300:]}"
    else:
        bad = [x for x in ("pac", "ann", "amb", "bwt", "sa")
               if open(os.path.join(d, "o." + x), "rb").read() != open(os.path.join(d, "e." + x), "rb").read()]
        ok, msg = not bad, "differ: " + ",".join(bad)
    if ok:
        shutil.rmtree(d, ignore_errors=True)
    return seed, ok, "" if ok else msg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--out", default="/tmp/b2fuzz_index")
    ap.add_argument("--engine", choices=["hostsim", "gpu"], default="hostsim")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    n_bad = 0
    with ProcessPoolExecutor(a.jobs) as ex:
        for seed, ok, msg in ex.map(trial, [(s, a.out, a.engine) for s in range(a.first, a.first + a.count)]):
            if not ok:
                n_bad += 1
                print("MISMATCH", seed, msg, flush=True)
    print(f"index trials {a.count}: {a.count - n_bad} identical, {n_bad} findings")


if __name__ == "__main__":
    main()
