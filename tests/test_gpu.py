"""GPU parity tests (run on the B200 box): the CUDA build, called through the C-ABI / CLI twin,
must reproduce the reference's SAM bit for bit (the @PG line excluded)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
ROOT = pu.ROOT
GOLD = os.path.join(ROOT, "tests", "golden", "tiny")
LIB = os.path.join(ROOT, "sparkbwa_b200", "libbwa.so")


def _lib():
    lib = ctypes.CDLL(LIB)
    lib.b200bwa_main_to.restype = ctypes.c_int
    lib.b200bwa_last_error.restype = ctypes.c_char_p
    lib.b200bwa_index_load.restype = ctypes.c_void_p
    lib.b200bwa_index_load.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.b200bwa_index_free.argtypes = [ctypes.c_void_p]
    lib.b200bwa_set_options.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.b200bwa_align_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p,
                                        ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_size_t)]
    lib.b200bwa_free.argtypes = [ctypes.c_void_p]
    return lib


def _main_to(lib, args, out):
    arr = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
    return lib.b200bwa_main_to(len(args), arr, out.encode())


@pytest.mark.parametrize("args,gold", [
    (["ref.fa", "se.fq"], "se.sam"),
    (["-K", "20000", "ref.fa", "pe_1.fq", "pe_2.fq"], "pe.sam"),
    (["-K", "20000", "-M", "-Y", "-R", r"@RG\tID:g1\tSM:s1", "-C", "ref.fa", "pe_1.fq", "pe_2.fq"], "pe_opts.sam"),
])
def test_cabi_matches_golden(tmp_path, args, gold):
    """In-process call through the C-ABI (what the JNI symbol forwards to) against committed vectors."""
    lib = _lib()
    out = str(tmp_path / "o.sam")
    full = ["bwa", "mem"] + [os.path.join(GOLD, a) if a.endswith((".fa", ".fq")) else a for a in args]
    rc = _main_to(lib, full, out)
    assert rc == 0, lib.b200bwa_last_error()
    assert pu.sam_body(out) == pu.sam_body(os.path.join(GOLD, gold))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(pu.CASES))
def test_cli_matches_oracle(workdir, name):
    c = pu.make_case(workdir, name)
    ref = pu.oracle_sam(c)
    our = pu.engine_sam(c, pu.CLI)
    assert not pu.first_diff(ref, our)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("extra", [["-a"], ["-k", "15", "-w", "40", "-T", "20"], ["-A", "2"], ["-S"], ["-P"],
                                   ["-I", "400,40"], ["-O", "4,8", "-E", "2,1"], ["-c", "20", "-y", "8"], ["-M", "-h", "3"]])
def test_cli_option_surface(workdir, extra):
    c = pu.make_case(workdir, "pe_small")
    ref = pu.oracle_sam(c, extra=extra)
    our = pu.engine_sam(c, pu.CLI, extra=extra)
    assert not pu.first_diff(ref, our)


def test_jni_symbol_end_to_end(tmp_path):
    """The exact SparkBWA argv (Bwa.java:289-390): bwa mem <-w tokens> -f <out> <idx> <fq1> <fq2>."""
    exe = str(tmp_path / "jni_fake")
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "jni_fake.c"), "-ldl"], check=True)
    out = str(tmp_path / "part0.sam")
    r = subprocess.run([exe, LIB, "bwa", "mem", "-K", "20000", "-f", out, os.path.join(GOLD, "ref.fa"),
                        os.path.join(GOLD, "pe_1.fq"), os.path.join(GOLD, "pe_2.fq")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b"rc=0" in r.stdout, r.stderr[-2000:]
    assert pu.sam_body(out) == pu.sam_body(os.path.join(GOLD, "pe.sam"))


def _load_fastq(path):
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        n = lines[i][1:]
        if len(n) > 2 and n[-2:-1] == b"/" and n[-1:].isdigit():
            n = n[:-2]
        names.append(n); seqs.append(lines[i + 1]); quals.append(lines[i + 3])
    return names, seqs, quals


def test_library_batch_api_host_buffers(tmp_path):
    """b200bwa_align_batch with reads in host memory == the file path, batch for batch."""
    lib = _lib()
    ix = lib.b200bwa_index_load(os.path.join(GOLD, "ref.fa").encode(), 0)
    assert ix, lib.b200bwa_last_error()
    assert lib.b200bwa_set_options(ix, b"") == 0
    n1, s1, q1 = _load_fastq(os.path.join(GOLD, "pe_1.fq"))
    n2, s2, q2 = _load_fastq(os.path.join(GOLD, "pe_2.fq"))
    names, seqs, quals = [], [], []
    for a, b, c, d, e, f in zip(n1, s1, q1, n2, s2, q2):
        names += [a, d]; seqs += [b, e]; quals += [c, f]
    # cut batches the way bseq_read does for -K 20000
    out = b""
    i = 0
    n_processed = 0
    while i < len(seqs):
        size = 0
        j = i
        while j < len(seqs):
            size += len(seqs[j]) + len(seqs[j + 1]); j += 2
            if size >= 20000:
                break
        off = np.zeros(j - i + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs[i:j]])
        sam = ctypes.c_void_p(); ln = ctypes.c_size_t()
        rc = lib.b200bwa_align_batch(ix, j - i, b"".join(seqs[i:j]), b"".join(quals[i:j]), off.ctypes.data,
                                     b"\0".join(names[i:j]) + b"\0", n_processed, 1, ctypes.byref(sam), ctypes.byref(ln))
        assert rc == 0, lib.b200bwa_last_error()
        out += ctypes.string_at(sam.value, ln.value)
        lib.b200bwa_free(sam)
        n_processed += j - i
        i = j
    lib.b200bwa_index_free(ix)
    gold = b"".join(l for l in pu.sam_body(os.path.join(GOLD, "pe.sam")) if not l.startswith(b"@"))
    assert out == gold


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_batch_size_independence_of_se(workdir):
    """Size-independent property: single-end output does not depend on -K (no batch-wide statistics)."""
    c = pu.make_case(workdir, "se100")
    a = pu.sam_digest(pu.engine_sam(c, pu.CLI, extra=["-K", "30000"]))
    b = pu.sam_digest(pu.engine_sam(c, pu.CLI, extra=["-K", "100000000"]))
    assert a == b


def test_repeat_call_same_process(tmp_path):
    """Two JNI-style calls in one process reuse the HBM-resident index and stay deterministic."""
    lib = _lib()
    outs = []
    for k in range(2):
        out = str(tmp_path / f"o{k}.sam")
        rc = _main_to(lib, ["bwa", "mem", os.path.join(GOLD, "ref.fa"), os.path.join(GOLD, "se.fq")], out)
        assert rc == 0
        outs.append(pu.sam_digest(out))
    assert outs[0] == outs[1]


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_smart_pairing(workdir):
    """-p: interleaved input, single-end reads mixed in, several batches (bwa/fastmap.c:64-83, bwa/bwa.c:77-93)."""
    c = pu.make_case(workdir, "pe_small")
    inter = pu.make_interleaved(c)
    c2 = dict(c, fq=[inter], opts=["-p", "-K", "30000"])
    ref = pu.oracle_sam(c2)
    our = pu.engine_sam(c2, pu.CLI)
    assert not pu.first_diff(ref, our)


def test_paired_names_must_match(workdir):
    """Out-of-step mate files end the run with the reference's message and status 1 (bwa/bwamem_pair.c:355,385), in the
    middle of a multi-batch job too; the output file is cut back to its header."""
    c = pu.make_case(workdir, "pe_small")
    bad = os.path.join(c["dir"], "r_2_shifted.fq")
    recs = open(c["fq"][1], "rb").read().split(b"\n")
    with open(bad, "wb") as f:   # drop record 300 of mate 2: with -K 30000 the first batches are still in step
        f.write(b"\n".join(recs[:1200] + recs[1204:]))
    out = os.path.join(c["dir"], "shifted.sam")
    e = subprocess.run([pu.CLI, "mem", "-K", "30000", "-f", out, c["ref"], c["fq"][0], bad], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert e.returncode == 1
    assert b"paired reads have different names" in e.stderr, e.stderr[-500:]   # (which batch reports first is a race)
    assert all(l.startswith(b"@") for l in open(out, "rb").read().splitlines())  # a failed run leaves the header alone


def _config_scale_index(tmp_path_factory, ref_len):
    """100 Mbp-class synthetic reference + index files written by the GPU builder, cached in /dev/shm."""
    import sys
    sys.path.insert(0, ROOT)
    from sparkbwa_b200 import synth, index
    base = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path_factory.mktemp("cfg"))
    d = os.path.join(base, f"b2cfg_L{ref_len}")
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "ref")
    ctg = synth.make_reference(ref_len, n_contigs=4, seed=1, repeat_frac=0.02)
    if not os.path.exists(prefix + ".sa"):
        index.build_index(ctg, prefix + ".tmp")   # the GPU builder; byte parity with `bwa index` has its own test
        for e in ("pac", "ann", "amb", "bwt", "sa"):
            os.replace(prefix + ".tmp." + e, prefix + "." + e)
    return d, prefix, ctg


def _compare_with_oracle(lib, d, tag, opts, prefix, fqs, min_lines):
    ours, ref = os.path.join(d, f"ours_{tag}.sam"), os.path.join(d, f"oracle_{tag}.sam")
    rc = _main_to(lib, ["bwa", "mem", "-v", "1"] + opts + [prefix] + fqs, ours)
    assert rc == 0, lib.b200bwa_last_error()
    with open(ref, "wb") as f:
        subprocess.run([pu.ORACLE, "mem", "-v", "1", "-t", str(os.cpu_count() or 1)] + opts + [prefix] + fqs, check=True,
                       stdout=f, stderr=subprocess.DEVNULL)
    diff = pu.first_diff(ref, ours)
    n_lines = len(pu.sam_body(ours))
    for p in (ours, ref):
        os.remove(p)
    assert not diff, diff
    assert n_lines >= min_lines


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_config1_scale_matches_oracle(tmp_path_factory):
    """BASELINE.json configs[1] at its full index size: 2x150 bp pairs against a 100 Mbp reference with -K 10000000
    (several batches, so the per-batch insert-size statistics matter).  B200BWA_FULL_PAIRS (default 150000; the
    round's own visit runs 1000000) sets the read count; the oracle is `bwa mem -t <host cores>` on the same files,
    the index files are written by the GPU builder (byte-identical to `bwa index`'s).  configs[4]'s batch-size
    sweep is covered on the same reads with -K 1000000 and -K 100000000 (different batch cuts, different pestat)."""
    from sparkbwa_b200 import synth
    pairs = int(os.environ.get("B200BWA_FULL_PAIRS", "150000"))
    ref_len = int(os.environ.get("B200BWA_FULL_REF", "100000000"))
    d, prefix, ctg = _config_scale_index(tmp_path_factory, ref_len)
    f1, f2 = os.path.join(d, "r_1.fq"), os.path.join(d, "r_2.fq")
    s1, s2 = [], []
    for a, b in synth.simulate_pairs_fast(ctg, pairs, 150, seed=4242, sub=0.01, indel=0.001):
        s1.append(a); s2.append(b)
    synth.write_fastq_fixed(f1, s1, 0, b"/1")
    synth.write_fastq_fixed(f2, s2, 0, b"/2")
    del ctg, s1, s2
    lib = _lib()
    try:
        _compare_with_oracle(lib, d, "K1e7", ["-K", "10000000"], prefix, [f1, f2], 2 * pairs)
        if pairs <= 200000:
            _compare_with_oracle(lib, d, "K1e6", ["-K", "1000000"], prefix, [f1, f2], 2 * pairs)
            _compare_with_oracle(lib, d, "K1e8", ["-K", "100000000"], prefix, [f1, f2], 2 * pairs)
    finally:
        for p in (f1, f2):
            os.remove(p)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_config3_scale_single_end_250bp(tmp_path_factory):
    """BASELINE.json configs[3]: single-end 250 bp reads at 3.5 % substitutions + 1.5 % indels (wide ksw_extend2 bands,
    band retries in mem_reg2aln) against the 100 Mbp index."""
    from sparkbwa_b200 import synth
    n = int(os.environ.get("B200BWA_FULL_SE", "30000"))
    d, prefix, ctg = _config_scale_index(tmp_path_factory, int(os.environ.get("B200BWA_FULL_REF", "100000000")))
    fq = os.path.join(d, "se250.fq")
    r1, _ = synth.simulate_reads(ctg, n=n, read_len=250, paired=False, seed=99, sub=0.035, indel=0.015)
    synth.write_fastq(fq, r1)
    del ctg, r1
    try:
        _compare_with_oracle(_lib(), d, "se250", [], prefix, [fq], n)
    finally:
        os.remove(fq)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_one_job_over_several_engines_matches_oracle(workdir):
    """ONE call, several engines: the -K batches of the job are dealt round-robin over the devices (each with its absolute
    n_processed), SAM blocks land in input order (bwa/kthread.c:92-102, bwa/fastmap.c:84-89).  On a multi-GPU box this
    uses every visible GPU (index fanned out device to device); on a single GPU the same device is listed twice, which
    gives two engines with their own index copy, streams and buffers -- the same dealing/ordering code."""
    import torch
    n = torch.cuda.device_count()
    devs = "all" if n >= 2 else "0,0"
    c = pu.make_case(workdir, "pe150")   # -K 300000: three batches
    ref = pu.oracle_sam(c)
    our = pu.engine_sam(c, pu.CLI, env={"B200BWA_DEVICES": devs, "B200BWA_TIMES": "1"})
    assert not pu.first_diff(ref, our)
    # smaller batches: 15 of them over the engines, and through a pipe (ordered writer instead of positioned writes)
    c2 = dict(c, opts=["-K", "60000"])
    ref2 = pu.oracle_sam(c2)
    for via_f in (True, False):
        our2 = pu.engine_sam(c2, pu.CLI, env={"B200BWA_DEVICES": devs}, via_f=via_f)
        assert not pu.first_diff(ref2, our2)


def test_device_ingest_is_the_path_taken(tmp_path):
    """Plain FASTQ files are packed on the device (k_ingest_*), not by host threads."""
    out = str(tmp_path / "o.sam")
    r = subprocess.run([pu.CLI, "mem", "-K", "20000", "-f", out, os.path.join(GOLD, "ref.fa"), os.path.join(GOLD, "pe_1.fq"),
                        os.path.join(GOLD, "pe_2.fq")], env=dict(os.environ, B200BWA_TIMES="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr[-2000:]
    assert b"device ingest" in r.stderr and b"host pack" not in r.stderr
    assert pu.sam_body(out) == pu.sam_body(os.path.join(GOLD, "pe.sam"))


@pytest.mark.skipif(not os.access(pu.ORACLE_BATCH, os.X_OK), reason="oracle/_ref/oracle_batch not built")
def test_absolute_read_index_and_int32_wrap(workdir):
    """SURVEY.md H9 on the device: hash tie-breaks from the absolute read index, `id<<8` wrapping for pair indices >= 2^23
    (what a 20 M-pair job reaches) -- against the reference's own mem_process_seqs called with that n_processed."""
    c = pu.make_case(workdir, "pe_repeat")
    assert pu.check_first_read_offsets(c, pu.CLI) > 100
    c = pu.make_case(workdir, "pe_small")
    assert pu.check_first_read_offsets(c, pu.CLI) > 10


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_partitions_map_and_reduce_like_sparkbwa(workdir, tmp_path):
    """SparkBWA without Spark (BwaInterpreter.java:301-376): the reads split into partitions, one `mem` call per
    partition, the reducer concatenating the partitions' SAM files (header of partition 0 only).  Expected: the
    reference's `bwa mem` run on each partition file, merged the same way."""
    c = pu.make_case(workdir, "pe_small")
    parts = 3

    def split(path, tag):
        recs = open(path, "rb").read().split(b"\n")
        n = (len(recs) - 1) // 4
        out = []
        for p in range(parts):
            f = str(tmp_path / f"RDD{p}_{tag}")
            with open(f, "wb") as fo:   # the shape BwaPairedAlignment writes: a blank line after every record
                for k in range(p * n // parts, (p + 1) * n // parts):
                    fo.write(b"\n".join(recs[4 * k:4 * k + 4]) + b"\n\n")
            out.append(f)
        return out
    f1, f2 = split(c["fq"][0], 1), split(c["fq"][1], 2)
    want = b""
    for p in range(parts):
        r = subprocess.run([pu.ORACLE, "mem", "-M", c["ref"], f1[p], f2[p]], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
        want += b"".join(l for l in r.stdout.splitlines(True) if not l.startswith(b"@PG") and (p == 0 or not l.startswith(b"@")))
    lib = _lib()
    opts = (ctypes.c_char_p * 1)(b"-M")
    a1 = (ctypes.c_char_p * parts)(*[x.encode() for x in f1])
    a2 = (ctypes.c_char_p * parts)(*[x.encode() for x in f2])
    od = tmp_path / "out"
    od.mkdir()
    rc = lib.b200bwa_run_partitions(1, opts, c["ref"].encode(), parts, a1, a2, str(od).encode(), b"SparkBWA-app1", 1)
    assert rc == 0, lib.b200bwa_last_error()
    got = b"".join(l for l in open(od / "FullOutput.sam", "rb") if not l.startswith(b"@PG"))
    assert got == want
    assert sorted(os.listdir(od)) == ["FullOutput.sam"]


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["pe_repeat", "pe250", "se250", "pe_alt"])
def test_in_kernel_dp_paths_match_oracle(workdir, name):
    """The DP results that the consumers find missing are computed in place by the lane-group forms -- single lanes in the
    per-read extension pass (b2_extend2_coop / b2_global2_coop with a one-lane group), 4-lane groups in the
    finalisation kernel (b2_global2_*, and b2_align2_g: ksw_u8 / ksw_i16 with the SSE lanes dealt over 4 lanes).
    With every run-ahead stage switched off ALL of the DP goes through those paths; the SAM must not change."""
    c = pu.make_case(workdir, name)
    ref = pu.oracle_sam(c)
    env = {"B200BWA_NO_XEXT": "1", "B200BWA_NO_EXT_CHAIN": "1", "B200BWA_NO_GALN": "1", "B200BWA_NO_RESC_PLAN": "1"}
    our = pu.engine_sam(c, pu.CLI, env=env)
    assert not pu.first_diff(ref, our)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["pe_small", "pe_repeat", "se_long900"])
def test_capacity_retry_paths_on_device(workdir, name):
    """Deliberately tiny per-lane regions, outlier-pool slots, interval tables and SAM slots: on the device too every
    stage must notice the overflow, run the task again in a pool slot or grow the buffer and re-run, and still produce
    the reference's SAM (the host-build twin is test_capacity_retry_paths; round 2's only device fault sat on such a path)."""
    c = pu.make_case(workdir, name)
    ref = pu.oracle_sam(c)
    env = {"B200BWA_SCRATCH": "1024", "B200BWA_BIG_SLOT": "4096", "B200BWA_BIG_POOL": str(1 << 20), "B200BWA_CAP_INTV": "2",
           "B200BWA_SLOT": "64", "B200BWA_WORKERS": "2"}
    assert not pu.first_diff(ref, pu.engine_sam(c, pu.CLI, env=env))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("env", [{"B200BWA_FINAL_HEAVY": "0"}, {"B200BWA_FINAL_HEAVY": "2"}, {"B200BWA_FINAL_HEAVY": "1000"},
                                 {"B200BWA_FINAL_FAST": "0"}, {"B200BWA_FINAL_FAST": "64"}],
                         ids=["one-launch", "every-pair-a-warp", "no-pair-listed", "no-shared-arena", "tiny-shared-arena"])
def test_pair_finalisation_split_does_not_change_the_sam(workdir, env):
    """Pairs with many regions are finalised by k_final_pe_heavy (a warp each), the others by k_final_pe (4-lane groups);
    which pair goes where, and whether the read copy / reference window live in shared memory, must not show in the
    output: repeat-rich pairs (dozens of regions, rescue, XA), 250 bp pairs and ALT contigs against the reference."""
    for name in ("pe_repeat", "pe250", "pe_alt"):
        c = pu.make_case(workdir, name)
        assert not pu.first_diff(pu.oracle_sam(c), pu.engine_sam(c, pu.CLI, env=env)), (name, env)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_sa_lookup_forms(workdir):
    """k_sa (lookups described by k_sa_plan, next lookup prefetched stage by stage) and k_sa_search (no tables) leave the
    same seeds: repeat-rich pairs, where a read has thousands of lookups, against the reference."""
    c = pu.make_case(workdir, "pe_repeat")
    ref = pu.oracle_sam(c)
    assert not pu.first_diff(ref, pu.engine_sam(c, pu.CLI, env={"B200BWA_NO_SA_PLAN": "1"}))
    assert not pu.first_diff(ref, pu.engine_sam(c, pu.CLI))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_index_builder_matches_bwa_index(workdir):
    """`b200bwa index` (suffix sort on the GPU) writes the files of `bwa index`, byte for byte: the small adversarial
    cases, a 5 Mbp genome with ambiguous bases (BASELINE.json configs[0]'s reference size) and -- with
    B200BWA_FULL_INDEX=1, about two minutes of single-threaded `bwa index` -- the 100 Mbp reference of configs[1]."""
    from sparkbwa_b200 import synth
    d = os.path.join(workdir, "index")
    for tag, ctg, with_n, env in pu.index_cases():
        pu.check_index_builder(pu.CLI, d, tag, ctg, with_n, env)
    pu.check_index_builder(pu.CLI, d, "g5M", synth.make_reference(5_000_000, n_contigs=3, seed=11, repeat_frac=0.03), True, {})
    pu.check_index_builder(pu.CLI, d, "g5M_groups", synth.make_reference(5_000_000, n_contigs=3, seed=12, repeat_frac=0.03), False,
                           {"B200BWA_INDEX_GROUP": "1500000"})
    if os.environ.get("B200BWA_FULL_INDEX") == "1":
        pu.check_index_builder(pu.CLI, d, "g100M", synth.make_reference(100_000_000, n_contigs=4, seed=1, repeat_frac=0.02), True, {})


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_concurrent_calls_in_one_process(workdir, tmp_path):
    """Three threads of one process call the C-ABI entry point at the same time (what `--executor-cores 3` does to the
    JNI symbol); engines are leased per call, every output equals the reference's."""
    import threading
    c = pu.make_case(workdir, "pe150")
    ref = pu.oracle_sam(c)
    lib = _lib()
    rcs = {}

    def call(k):
        rcs[k] = _main_to(lib, ["bwa", "mem", "-v", "1"] + c["opts"] + [c["ref"]] + c["fq"], str(tmp_path / f"o{k}.sam"))
    th = [threading.Thread(target=call, args=(k,)) for k in range(3)]
    [t.start() for t in th]; [t.join() for t in th]
    assert rcs == {0: 0, 1: 0, 2: 0}
    for k in range(3):
        assert not pu.first_diff(ref, str(tmp_path / f"o{k}.sam"))
