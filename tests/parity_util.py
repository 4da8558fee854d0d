"""Shared helpers of the parity tests: seeded workloads, the oracle binary, SAM comparison.

The oracle (`oracle/_ref/bwa`, the unmodified reference compiled by oracle/Makefile) is test
infrastructure only: it builds the index files and produces the SAM the engine must reproduce.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparkbwa_b200 import synth  # noqa: E402

ORACLE = os.path.join(ROOT, "oracle", "_ref", "bwa")
CLI = os.path.join(ROOT, "sparkbwa_b200", "b200bwa")
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "_build", "b200bwa_hostsim")
ORACLE_BATCH = os.path.join(ROOT, "oracle", "_ref", "oracle_batch")

# name -> (reference kwargs, reads kwargs, extra bwa-mem options)
CASES = {
    # BASELINE.json configs[0] at its stated size: 10 k x 100 bp single-end reads, 2 % error, vs a 5 Mbp reference
    "cfg0_se100_5Mbp": (dict(total_len=5_000_000, n_contigs=1, seed=1, repeat_frac=0.02),
                        dict(n=10000, read_len=100, paired=False, seed=21, sub=0.02, indel=0.0), []),
    # config 1 shape (scaled): single-end 100 bp, 2 % error
    "se100": (dict(total_len=1_000_000, n_contigs=3, seed=1, repeat_frac=0.05),
              dict(n=2000, read_len=100, paired=False, seed=3, sub=0.02, indel=0.002, n_frac=0.002, junk_frac=0.01), []),
    # config 2 shape (scaled): paired-end 2x150 bp, 1 % error, several -K batches
    "pe150": (dict(total_len=1_000_000, n_contigs=3, seed=1, repeat_frac=0.05),
              dict(n=3000, read_len=150, paired=True, seed=5, sub=0.01, indel=0.001, junk_frac=0.01), ["-K", "300000"]),
    # config 4 shape: single-end 250 bp, 3.5 % substitutions + 1.5 % indels (wide bands)
    "se250": (dict(total_len=600_000, n_contigs=2, seed=2, repeat_frac=0.03),
              dict(n=800, read_len=250, paired=False, seed=9, sub=0.035, indel=0.015), []),
    # noisy mate 2: forces mate rescue (ksw_u8 path)
    "pe_rescue": (dict(total_len=500_000, n_contigs=2, seed=4, repeat_frac=0.03),
                  dict(n=1500, read_len=150, paired=True, seed=11, sub=0.01, indel=0.001, mate2_sub=0.10), []),
    # repeat-rich reference: deep B-trees, max_occ sampling, many rescues, XA tags
    "pe_repeat": (dict(total_len=300_000, n_contigs=2, seed=6, repeat_frac=0.60, rep_len=(300, 3000), rep_copies=(5, 60), rep_div=0.04),
                  dict(n=1200, read_len=150, paired=True, seed=13, sub=0.01, indel=0.001), []),
    # 250 bp pairs -> 16-bit local SW in mate rescue
    "pe250": (dict(total_len=400_000, n_contigs=1, seed=8, repeat_frac=0.10),
              dict(n=600, read_len=250, paired=True, seed=15, sub=0.02, indel=0.004, mate2_sub=0.08, ins_mean=600.0, ins_sd=60.0), []),
    # chimeric reads -> supplementary alignments (SA tag, 0x800, hard clips)
    "pe_chimeric": (dict(total_len=400_000, n_contigs=3, seed=31, repeat_frac=0.04),
                    dict(n=800, read_len=150, paired=True, seed=33, sub=0.01, indel=0.001, chimeric_frac=0.25), []),
    # short and ragged reads: below min_seed_len, empty alignments, length 1
    "se_short": (dict(total_len=200_000, n_contigs=1, seed=35, repeat_frac=0.02),
                 dict(n=600, read_len=60, paired=False, seed=37, sub=0.03, indel=0.003, n_frac=0.02, ragged=True), []),
    # long reads: mem_flt_chained_seeds re-scores every seed (reads >= 725 bp); repeats + chimeras make it drop some
    "se_long900": (dict(total_len=300_000, n_contigs=2, seed=61, repeat_frac=0.5, rep_len=(100, 800), rep_copies=(5, 40), rep_div=0.08),
                   dict(n=150, read_len=900, paired=False, seed=63, sub=0.08, indel=0.03, junk_frac=0.05, chimeric_frac=0.3), []),
    # the pacbio preset (-x pacbio: a=b=o=e=1, -k17 -W40 -r10 -L0) on 1.5 kb reads at 14 % error
    "se_pacbio1500": (dict(total_len=300_000, n_contigs=2, seed=71, repeat_frac=0.5, rep_len=(100, 800), rep_copies=(5, 40), rep_div=0.08),
                      dict(n=100, read_len=1500, paired=False, seed=73, sub=0.10, indel=0.04, chimeric_frac=0.3), ["-x", "pacbio"]),
    # ALT contigs listed in ref.fa.alt: ALT-aware primary marking, ALT supplementary records, XA with max_XA_hits_alt
    "pe_alt": (dict(total_len=300_000, n_contigs=2, seed=81, repeat_frac=0.05, alt=dict(n_alt=8, alt_len=(1200, 4000), div=0.02, seed=83)),
               dict(n=1500, read_len=150, paired=True, seed=85, sub=0.01, indel=0.001, chimeric_frac=0.05), []),
    "se_alt": (dict(total_len=200_000, n_contigs=2, seed=87, repeat_frac=0.10, rep_len=(200, 1500), rep_copies=(2, 8),
                    alt=dict(n_alt=10, alt_len=(800, 3000), div=0.03, seed=89)),
               dict(n=1500, read_len=120, paired=False, seed=91, sub=0.015, indel=0.002), ["-a"]),
    # the other two presets of bwa/fastmap.c:288-317
    "se_ont2d1200": (dict(total_len=300_000, n_contigs=2, seed=93, repeat_frac=0.3, rep_len=(100, 800), rep_copies=(3, 20), rep_div=0.08),
                     dict(n=80, read_len=1200, paired=False, seed=95, sub=0.10, indel=0.05, chimeric_frac=0.2), ["-x", "ont2d"]),
    "se_intractg2000": (dict(total_len=300_000, n_contigs=2, seed=97, repeat_frac=0.3, rep_len=(100, 800), rep_copies=(3, 20), rep_div=0.05),
                        dict(n=80, read_len=2000, paired=False, seed=99, sub=0.01, indel=0.002, chimeric_frac=0.3), ["-x", "intractg"]),
    # found by scripts/fuzz_parity.py: 2.5 kb reads seeded with -k 5 on a repeat-rich reference have more than 8 192 seed
    # intervals per read (re-seeding, bwamem.c:131-144), which the interval table once refused to grow to
    "se_k5_long2500": (dict(total_len=80_000, n_contigs=2, seed=7005482, repeat_frac=0.8, rep_len=(60, 250), rep_copies=(2, 10), rep_div=0.02),
                       dict(n=20, read_len=2500, paired=False, seed=7005484, sub=0.2, indel=0.0, junk_frac=0.1, chimeric_frac=0.6),
                       ["-a", "-S", "-k", "5"]),
    # reads beyond 65 535 bases (query positions no longer fit 16 bits; lengths beyond the first 2^16 log-table entries)
    "se_ont2d_70k": (dict(total_len=600_000, n_contigs=2, seed=5, repeat_frac=0.1, rep_len=(100, 800), rep_copies=(2, 10), rep_div=0.05),
                     dict(n=3, read_len=70_000, paired=False, seed=9, sub=0.08, indel=0.03, chimeric_frac=0.5), ["-x", "ont2d"]),
    "pe_small": (dict(total_len=200_000, n_contigs=2, seed=41, repeat_frac=0.20, rep_len=(200, 1500), rep_copies=(2, 12)),
                 dict(n=400, read_len=125, paired=True, seed=43, sub=0.015, indel=0.002, mate2_sub=0.05, chimeric_frac=0.05), []),
}


def have_oracle() -> bool:
    return os.access(ORACLE, os.X_OK)


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def make_case(workdir: str, name: str):
    """Generate reference + reads + index (+ oracle SAM).  Returns dict of paths and options."""
    ref_kw, reads_kw, opts = CASES[name]
    ref_kw = dict(ref_kw)
    alt_kw = ref_kw.pop("alt", None)

    def reference():
        ctg = synth.make_reference(**ref_kw)
        alt = []
        if alt_kw:
            ctg, alt = synth.add_alt_contigs(ctg, **alt_kw)
        return ctg, alt
    d = os.path.join(workdir, name)
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    out = {"dir": d, "ref": fa, "opts": list(opts), "paired": reads_kw["paired"]}
    ctg = None
    if not os.path.exists(fa + ".bwt"):
        ctg, alt = reference()
        synth.write_fasta(fa, ctg)
        if alt:  # bwakit's .alt is a SAM file of the ALT contigs against the primary assembly; only column 1 is read
            with open(fa + ".alt", "w") as f:
                f.write("@SQ\tSN:chr1\tLN:1\n")
                for nm in alt:
                    f.write(f"{nm}\t0\tchr1\t1\t60\t100M\t*\t0\t0\t*\t*\n")
        run([ORACLE, "index", fa])
    fq1 = os.path.join(d, "r_1.fq")
    fq2 = os.path.join(d, "r_2.fq")
    if not os.path.exists(fq1):
        if ctg is None:
            ctg, _ = reference()
        r1, r2 = synth.simulate_reads(ctg, **reads_kw)
        synth.write_fastq(fq1, r1)
        if r2 is not None:
            synth.write_fastq(fq2, r2)
    out["fq"] = [fq1] + ([fq2] if reads_kw["paired"] else [])
    return out


def oracle_sam(case, extra=()):
    p = os.path.join(case["dir"], "oracle.sam")
    cmd = [ORACLE, "mem"] + case["opts"] + list(extra) + [case["ref"]] + case["fq"]
    with open(p, "wb") as f:
        subprocess.run(cmd, check=True, stdout=f, stderr=subprocess.DEVNULL)
    return p


def engine_sam(case, binary, extra=(), env=None, via_f=True):
    p = os.path.join(case["dir"], "engine.sam")
    cmd = [binary, "mem"] + case["opts"] + list(extra) + (["-f", p] if via_f else []) + [case["ref"]] + case["fq"]
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    if r.returncode != 0:
        raise RuntimeError(f"engine failed rc={r.returncode}: {r.stderr.decode()[-2000:]}")
    if not via_f:
        with open(p, "wb") as f:
            f.write(r.stdout)
    return p


def sam_body(path):
    with open(path, "rb") as f:
        return [l for l in f if not l.startswith(b"@PG")]


def sam_digest(path):
    h = hashlib.md5()
    for l in sam_body(path):
        h.update(l)
    return h.hexdigest()


def first_diff(a, b, limit=3):
    la, lb = sam_body(a), sam_body(b)
    out = []
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            out.append((i, x.decode(errors="replace")[:300], y.decode(errors="replace")[:300]))
            if len(out) >= limit:
                break
    if len(la) != len(lb):
        out.append((min(len(la), len(lb)), f"len {len(la)}", f"len {len(lb)}"))
    return out


HOSTSIM_ENV = {"B200BWA_LANE_GRID": "1", "B200BWA_LANE_BLOCK": "1", "B200BWA_SEED_GRID": "1",
               "B200BWA_SEED_BLOCK": "1", "B200BWA_SCRATCH": str(4 << 20)}


def make_interleaved(case, drop_every=7):
    """Interleaved FASTQ for smart pairing (-p): the two mate files merged pair by pair, with mate 2 of every
    `drop_every`-th pair and mate 1 of every (drop_every+4)-th pair removed so that single-end reads occur too."""
    out = os.path.join(case["dir"], "inter.fq")
    with open(case["fq"][0], "rb") as f1, open(case["fq"][1], "rb") as f2, open(out, "wb") as fo:
        k = 0
        while True:
            a = [f1.readline() for _ in range(4)]
            b = [f2.readline() for _ in range(4)]
            if not a[0] or not b[0]:
                break
            if k % (drop_every + 4) != 3:
                fo.writelines(a)
            if k % drop_every != 5:
                fo.writelines(b)
            k += 1
    return out


def check_first_read_offsets(case, binary, env=None, offsets=(0, (1 << 24) + 2 * 4321, (1 << 25) - 2, (1 << 31) + 6)):
    """The absolute read index (n_processed, bwa/fastmap.c:85) feeds hash_64 tie-breaks (bwa/bwamem.c:527,1162,1167)
    and, as a 32-bit `id<<8`, mem_pair (bwa/bwamem_pair.c:222), where it wraps for pair indices >= 2^23.
    oracle/_ref/oracle_batch runs the reference's mem_process_seqs on the case's reads as ONE batch with a chosen
    n_processed; the engine gets the same offset through B200BWA_FIRST_READ.  Returns how many SAM lines of the
    oracle change with the offset (the check is only meaningful if some do)."""
    base, moved = None, 0
    for n0 in offsets:
        r = subprocess.run([ORACLE_BATCH, case["ref"], str(n0)] + case["fq"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
        ref = r.stdout.splitlines(True)
        our = engine_sam(case, binary, extra=["-K", "2000000000"], env=dict(env or {}, B200BWA_FIRST_READ=str(n0)))
        ours = [l for l in sam_body(our) if not l.startswith(b"@")]
        assert ours == ref, f"SAM differs from the reference's mem_process_seqs at n_processed={n0}"
        if base is None:
            base = ref
        else:
            moved = max(moved, sum(1 for a, b in zip(ref, base) if a != b))
    return moved


def index_cases():
    """(tag, contigs, plant ambiguous bases?, builder env) for the index-construction parity tests: random genomes with
    exact repeats (many refinement rounds), ambiguous bases (lrand48 replay, .amb holes incl. mixed letters and lower
    case), several suffix groups, every text length around the 4 / 32 / 128 packing boundaries, homopolymers, tandem
    repeats and a reverse-complement palindrome (forward and reverse strand suffixes tie until the very end)."""
    import numpy as np
    rng = np.random.default_rng(9)
    out = [("rep", synth.make_reference(5000, n_contigs=2, seed=3, repeat_frac=0.3, rep_len=(100, 600), rep_copies=(2, 5), rep_div=0.0), False, {}),
           ("mid", synth.make_reference(200_000, n_contigs=3, seed=4, repeat_frac=0.2, rep_len=(300, 3000), rep_copies=(2, 10), rep_div=0.01), False, {}),
           ("groups", synth.make_reference(120_000, n_contigs=3, seed=5, repeat_frac=0.2, rep_len=(300, 3000), rep_copies=(2, 10), rep_div=0.01), True,
            {"B200BWA_INDEX_GROUP": "30000"}),
           ("withN", synth.make_reference(150_000, n_contigs=2, seed=6, repeat_frac=0.1), True, {})]
    for L in (1, 2, 3, 5, 16, 31, 33, 64, 127, 129, 1000):
        out.append((f"len{L}", [("c", rng.integers(0, 4, size=L, dtype=np.uint8))], False, {}))
    tand = np.tile(rng.integers(0, 4, size=7, dtype=np.uint8), 400)
    pal = rng.integers(0, 4, size=800, dtype=np.uint8)
    out += [("polyA", [("a", np.zeros(3000, dtype=np.uint8)), ("b", rng.integers(0, 4, size=500, dtype=np.uint8))], False, {}),
            ("polyTA", [("a", np.concatenate([np.full(700, 3, dtype=np.uint8), np.zeros(900, dtype=np.uint8)]))], False, {}),
            ("tandem", [("t", np.concatenate([tand, rng.integers(0, 4, size=100, dtype=np.uint8), tand[:1500]]))], False, {}),
            ("palindrome", [("p desc of p", np.concatenate([pal, (3 - pal[::-1]).astype(np.uint8)]))], False, {}),
            ("polyA_groups", [("a", np.zeros(3000, dtype=np.uint8)), ("b", rng.integers(0, 4, size=5000, dtype=np.uint8))], False,
             {"B200BWA_INDEX_GROUP": "1000"})]
    return out


def plant_ambiguous(fa, seed=5, runs=40):
    import numpy as np
    s = bytearray(open(fa, "rb").read())
    rng = np.random.default_rng(seed)
    for k in range(runs):
        i = int(rng.integers(0, max(1, len(s) - 200)))
        for j in range(i, min(len(s), i + int(rng.integers(1, 120)))):
            if s[j] in b"ACGT":
                s[j] = ord("N") if k % 3 else ord("R") if k % 2 else ord("n")
    open(fa, "wb").write(bytes(s))


def check_index_builder(binary, d, tag, contigs, with_n, env):
    """`<binary> index` must write the five files `bwa index` writes, byte for byte."""
    import filecmp
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, tag + ".fa")
    synth.write_fasta(fa, contigs)
    if with_n:
        plant_ambiguous(fa)
    run([ORACLE, "index", "-p", os.path.join(d, tag + ".ref"), fa])
    r = subprocess.run([binary, "index", "-p", os.path.join(d, tag + ".our"), fa], env=dict(os.environ, **env),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-1500:]
    for e in ("pac", "ann", "amb", "bwt", "sa"):
        assert filecmp.cmp(os.path.join(d, tag + ".ref." + e), os.path.join(d, tag + ".our." + e), shallow=False), (tag, e)
