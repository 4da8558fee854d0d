"""CPU-side tests (no GPU): oracle pinned to the golden vectors, host logic, ABI surface.

The kernel bodies are additionally compiled for the host by tests/hostsim (test scaffolding, never
part of libbwa.so) and must reproduce the reference SAM bit for bit; this checks the control logic
here, the `-m gpu` tests check the CUDA build on a B200.
"""
import ctypes
import os
import re
import subprocess

import pytest

import parity_util as pu

ROOT = pu.ROOT
GOLD = os.path.join(ROOT, "tests", "golden", "tiny")
LIB = os.path.join(ROOT, "sparkbwa_b200", "libbwa.so")


@pytest.fixture(scope="session")
def hostsim():
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.run(["make", "-C", d, "-j4"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return pu.HOSTSIM


# ---- the oracle is pinned to the committed golden vectors -------------------------------------
@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,gold", [
    (["ref.fa", "se.fq"], "se.sam"),
    (["-K", "20000", "ref.fa", "pe_1.fq", "pe_2.fq"], "pe.sam"),
])
def test_oracle_reproduces_golden(tmp_path, args, gold):
    r = subprocess.run([pu.ORACLE, "mem"] + args, cwd=GOLD, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    body = [l for l in r.stdout.splitlines(True) if not l.startswith(b"@PG")]
    assert body == pu.sam_body(os.path.join(GOLD, gold))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_oracle_thread_independence(tmp_path):
    """Output depends on -K but not on -t: the property multi-GPU sharding relies on."""
    outs = []
    for t in ("1", "3"):
        r = subprocess.run([pu.ORACLE, "mem", "-t", t, "-K", "20000", "ref.fa", "pe_1.fq", "pe_2.fq"], cwd=GOLD,
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
        outs.append([l for l in r.stdout.splitlines(True) if not l.startswith(b"@PG")])
    assert outs[0] == outs[1]


# ---- kernel bodies on the host vs golden / oracle ---------------------------------------------
@pytest.mark.parametrize("args,gold", [
    (["ref.fa", "se.fq"], "se.sam"),
    (["-K", "20000", "ref.fa", "pe_1.fq", "pe_2.fq"], "pe.sam"),
    (["-K", "20000", "-M", "-Y", "-R", r"@RG\tID:g1\tSM:s1", "-C", "ref.fa", "pe_1.fq", "pe_2.fq"], "pe_opts.sam"),
])
def test_hostsim_matches_golden(hostsim, tmp_path, args, gold):
    out = str(tmp_path / "o.sam")
    r = subprocess.run([hostsim, "mem"] + args, cwd=GOLD, stdout=open(out, "wb"), stderr=subprocess.PIPE,
                       env=dict(os.environ, **pu.HOSTSIM_ENV))
    assert r.returncode == 0, r.stderr.decode()[-1000:]
    assert pu.sam_body(out) == pu.sam_body(os.path.join(GOLD, gold))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["se100", "pe150", "se250", "pe_rescue", "pe_repeat", "pe250", "pe_chimeric", "se_short", "se_long900",
                                  "se_pacbio1500", "pe_alt", "se_alt", "se_ont2d1200", "se_intractg2000", "se_k5_long2500", "se_ont2d_70k"])
def test_hostsim_matches_oracle(hostsim, workdir, name):
    c = pu.make_case(workdir, name)
    ref = pu.oracle_sam(c)
    our = pu.engine_sam(c, hostsim, env=pu.HOSTSIM_ENV)
    assert not pu.first_diff(ref, our)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
@pytest.mark.parametrize("extra", [["-a"], ["-k", "15", "-w", "40", "-T", "20"], ["-A", "2"], ["-S"], ["-P"],
                                   ["-I", "400,40"], ["-U", "9", "-L", "3,7"], ["-O", "4,8", "-E", "2,1"],
                                   ["-c", "20", "-y", "8"], ["-m", "3", "-D", "0.3"], ["-M", "-h", "3"]])
def test_hostsim_option_surface(hostsim, workdir, extra):
    c = pu.make_case(workdir, "pe_small")
    ref = pu.oracle_sam(c, extra=extra)
    our = pu.engine_sam(c, hostsim, extra=extra, env=pu.HOSTSIM_ENV)
    assert not pu.first_diff(ref, our)


# ---- ABI surface -------------------------------------------------------------------------------
def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200bwa.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200bwa_[a-z_]+)\s*\(", txt)))


def test_library_exports_declared_symbols():
    assert os.path.exists(LIB), "libbwa.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(LIB)
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert hasattr(lib, "Java_com_github_sparkbwa_BwaJni_bwa_1jni")


def test_error_behaviour_without_exit(tmp_path):
    """Nothing may exit the process (a JVM in production): failures come back as return codes."""
    lib = ctypes.CDLL(LIB)
    lib.b200bwa_main_to.restype = ctypes.c_int
    lib.b200bwa_last_error.restype = ctypes.c_char_p

    def call(args, out=None):
        arr = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
        return lib.b200bwa_main_to(len(args), arr, out.encode() if out else None)

    assert call(["bwa"]) == 1
    assert call(["bwa", "aln", "x.fa", "r.fq"]) == 1                # `mem` and `index` are provided, nothing else
    assert b"unrecognized command" in lib.b200bwa_last_error()
    assert call(["bwa", "index", "/nonexistent/x.fa"]) == 1
    assert b"fail to open" in lib.b200bwa_last_error()
    assert call(["bwa", "mem"]) == 1                                 # usage
    assert call(["bwa", "mem", "-Z", "a", "b"]) == 1                 # unknown option
    assert call(["bwa", "mem", "-R", "RG\\tID:x", "a", "b"]) == 1    # bad read group
    assert call(["bwa", "mem", "/nonexistent/idx", os.path.join(GOLD, "se.fq")], str(tmp_path / "o.sam")) == 1
    assert b"index" in lib.b200bwa_last_error()


def test_no_cpu_fallback_without_cuda(tmp_path):
    """With no CUDA device the aligning call must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    out = str(tmp_path / "o.sam")
    r = subprocess.run([pu.CLI, "mem", "-f", out, os.path.join(GOLD, "ref.fa"), os.path.join(GOLD, "se.fq")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    assert b"CUDA" in r.stderr or b"cuda" in r.stderr


def test_jni_symbol_through_fake_env(tmp_path):
    exe = str(tmp_path / "jni_fake")
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "jni_fake.c"), "-ldl"], check=True)
    r = subprocess.run([exe, LIB, "bwa", "aln", "x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b"rc=1 released=3" in r.stdout, (r.stdout, r.stderr)
    # "-f <out>" is consumed by the glue and not passed on as an aligner option
    r = subprocess.run([exe, LIB, "bwa", "mem", "-f", str(tmp_path / "x.sam"), "/nonexistent/idx", "r.fq"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b"rc=1 released=6" in r.stdout, (r.stdout, r.stderr)
    assert b"invalid option" not in r.stderr
    # a null String element must not reach strcmp(): rc 1, the strings acquired so far are released
    r = subprocess.run([exe, LIB, "bwa", "mem", "@NULL@", "r.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b"rc=1 released=2" in r.stdout, (r.stdout, r.stderr)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_capacity_retry_paths(hostsim, workdir):
    """Deliberately tiny interval / scratch / SAM-slot capacities: every stage must notice the overflow,
    grow the buffer, re-run (the kernels are idempotent) and still produce the reference's SAM."""
    c = pu.make_case(workdir, "pe_small")
    ref = pu.oracle_sam(c)
    env = dict(pu.HOSTSIM_ENV, B200BWA_SCRATCH="512", B200BWA_CAP_INTV="2", B200BWA_SLOT="64")
    our = pu.engine_sam(c, hostsim, env=env)
    assert not pu.first_diff(ref, our)
    # per-lane scratch beyond the budget: the stage re-runs with fewer lanes instead of giving up (wide-band long
    # reads must not abort a run the reference completes); a budget no single read fits fails loudly
    c = pu.make_case(workdir, "se_long900")
    ref = pu.oracle_sam(c)
    env = dict(pu.HOSTSIM_ENV, B200BWA_LANE_GRID="8", B200BWA_SCRATCH="512", B200BWA_MAX_SCRATCH=str(1 << 20), B200BWA_WORKERS="1")
    our = pu.engine_sam(c, hostsim, env=env)
    assert not pu.first_diff(ref, our)
    with pytest.raises(RuntimeError, match="scratch limit"):
        pu.engine_sam(c, hostsim, env=dict(env, B200BWA_MAX_SCRATCH="4096"))


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_pair_finalisation_split_host_logic(hostsim, workdir):
    """k_final_plan / k_final_pe_heavy / k_final_pe: whichever pairs are listed for the first launch (none, all, those with
    >= 2 or >= 4 regions), each pair is finalised exactly once and the SAM equals the reference's; a retry after a slot /
    scratch overflow in either launch re-runs both.  (Host build: one lane per group; the GPU test repeats this on device.)"""
    for name in ("pe_small", "pe_repeat"):
        c = pu.make_case(workdir, name)
        ref = pu.oracle_sam(c)
        for heavy in ("0", "2", "4", "1000"):
            env = dict(pu.HOSTSIM_ENV, B200BWA_FINAL_HEAVY=heavy)
            assert not pu.first_diff(ref, pu.engine_sam(c, hostsim, env=env)), (name, heavy)
        env = dict(pu.HOSTSIM_ENV, B200BWA_FINAL_HEAVY="2", B200BWA_SCRATCH="512", B200BWA_SLOT="64")
        assert not pu.first_diff(ref, pu.engine_sam(c, hostsim, env=env)), name


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_sa_lookup_forms(hostsim, workdir):
    """k_sa takes each lookup's (read, interval, rank) from k_sa_plan's tables; k_sa_search -- used when max_occ or the
    interval capacity does not fit 16 bits -- finds them by searching seed_off.  Same seeds either way."""
    c = pu.make_case(workdir, "pe_repeat")
    ref = pu.oracle_sam(c)
    assert not pu.first_diff(ref, pu.engine_sam(c, hostsim, env=dict(pu.HOSTSIM_ENV, B200BWA_NO_SA_PLAN="1")))
    c2 = dict(c, opts=c["opts"] + ["-c", "70000"])
    assert not pu.first_diff(pu.oracle_sam(c2), pu.engine_sam(c2, hostsim, env=pu.HOSTSIM_ENV))


def test_striped_sw_scan_equals_lazy_f(tmp_path):
    """The mate-rescue SW replaces the reference's lazy-F loop (ksw.c:176-188) by a max-plus carry scan;
    fuzz it against a literal emulation of that loop (byte and word mode, random scorings)."""
    exe = str(tmp_path / "sw_fuzz")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "sparkbwa_b200", "csrc"), "-I" + os.path.join(ROOT, "tests"),
                    "-I" + os.path.join(ROOT, "tests", "hostsim"), "-o", exe, os.path.join(ROOT, "tests", "sw_fuzz.cpp")], check=True)
    r = subprocess.run([exe, "12000", "777"], stdout=subprocess.PIPE, check=False)
    # second line: the generic-group form (groups of 1..16 lanes: what the device runs for groups smaller than the
    # 16 / 8 SSE lanes) against the scalar form, including -O x,0 (literal lazy-F loop) and the cell counts
    assert r.returncode == 0 and r.stdout.count(b"mismatches 0") == 2, r.stdout


def test_register_resident_coop_dp_equals_scalar(tmp_path):
    """The device runs ksw_extend2 / ksw_global2 as register-resident lane-group routines (b2_coopreg.cuh);
    the same templates, driven by a coroutine emulation of the lane group (simt_emu.h), are fuzzed against the
    scalar forms that the host checker build pins to the oracle (scores, end points, CIGARs, direction
    matrices and the algorithmic cell counts must all agree).  The same run checks the inter-task global alignment
    of b2_galn.cuh (one alignment per lane, band in registers by diagonal) against ksw_global2's scalar form."""
    exe = str(tmp_path / "coop_fuzz")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(ROOT, "sparkbwa_b200", "csrc"),
                    "-I" + os.path.join(ROOT, "tests"), "-I" + os.path.join(ROOT, "tests", "hostsim"), "-o", exe,
                    os.path.join(ROOT, "tests", "coop_fuzz.cpp")], check=True)
    r = subprocess.run([exe, "1200", "31337"], stdout=subprocess.PIPE, check=False)
    assert r.returncode == 0 and r.stdout.count(b"mismatches 0") == 4, r.stdout


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_damaged_index_files_are_refused(hostsim, workdir, tmp_path):
    """An index cut short or mixed up on its way to the executor's disk: the reference aborts (`inconsistent .ann and
    .amb files`, bwa/bntseq.c:146) or simply crashes (a truncated .bwt: SIGSEGV in the first rank query); inside a JVM
    neither may happen -- every such file set is refused with status 1 and a message before anything is uploaded."""
    import shutil
    c = pu.make_case(workdir, "pe_small")
    env = dict(os.environ, **pu.HOSTSIM_ENV)

    def attempt(damage):
        d = str(tmp_path / damage.__name__)
        os.makedirs(d)
        for e in ("", ".bwt", ".sa", ".pac", ".ann", ".amb"):
            shutil.copy(c["ref"] + e, os.path.join(d, "ref.fa" + e))
        damage(os.path.join(d, "ref.fa"))
        r = subprocess.run([hostsim, "mem", "-f", os.path.join(d, "out.sam"), os.path.join(d, "ref.fa"), c["fq"][0]],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert r.returncode == 1, (damage.__name__, r.returncode, r.stderr[-300:])
        assert b"[E::" in r.stderr, damage.__name__
        return r.stderr

    def short_bwt(p):
        with open(p + ".bwt", "r+b") as f:
            f.truncate(5000)

    def short_sa(p):
        with open(p + ".sa", "r+b") as f:
            f.truncate(os.path.getsize(p + ".sa") // 2)

    def short_pac(p):
        with open(p + ".pac", "r+b") as f:
            f.truncate(100)

    def foreign_amb(p):
        txt = open(p + ".amb").read().split("\n")
        txt[0] = "12345 " + txt[0].split(" ", 1)[1]
        open(p + ".amb", "w").write("\n".join(txt))

    def foreign_bwt(p):   # a .bwt/.sa pair of another (shorter) reference next to this .ann/.pac
        other = pu.make_case(workdir, "se250")
        shutil.copy(other["ref"] + ".bwt", p + ".bwt")
        shutil.copy(other["ref"] + ".sa", p + ".sa")

    def zero_sa_interval(p):
        with open(p + ".sa", "r+b") as f:
            f.seek(40)
            f.write(b"\0" * 8)

    assert b"bwt" in attempt(short_bwt)
    assert b".sa" in attempt(short_sa)
    assert b".pac" in attempt(short_pac)
    assert b"inconsistent .ann and .amb" in attempt(foreign_amb)
    assert b"inconsistent .ann and .bwt" in attempt(foreign_bwt)
    assert b"sampling interval" in attempt(zero_sa_interval)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_malformed_records_end_the_input_like_the_reference(hostsim, workdir):
    """A quality string longer than its sequence makes kseq_read return -2: bseq_read ends the batch there and the next
    attempt resumes at the next '@'.  The reference's pipeline has two threads taking turns at the reading step (one
    with -1) and each leaves at its first empty attempt (bwa/kthread.c:92-102, bwa/fastmap.c:44-47), so the input ends at
    the second (first) malformed record that opens an attempt, or at EOF.  Four bad records: after 7 reads, two in a row
    after 22 more (the second of them opens an attempt), one further on."""
    c = pu.make_case(workdir, "pe_small")
    recs = open(c["fq"][0], "rb").read().split(b"\n")
    bad = os.path.join(c["dir"], "midbad.fq")
    with open(bad, "wb") as f:
        for k in range(60):
            r = recs[4 * k:4 * k + 4]
            if k in (7, 30, 31, 45):
                r = r[:3] + [r[3] + b"IIIIII"]
            f.write(b"\n".join(r) + b"\n")
    n_lines = {}
    for opts in ([], ["-1"], ["-K", "1000"], ["-K", "1000", "-1"]):
        c1 = dict(c, fq=[bad], opts=opts)
        ref = pu.oracle_sam(c1)
        for env in (pu.HOSTSIM_ENV, dict(pu.HOSTSIM_ENV, B200BWA_NO_MMAP="1")):
            assert not pu.first_diff(ref, pu.engine_sam(c1, hostsim, env=env)), (opts, env)
        n_lines[" ".join(opts)] = len(set(l.split(b"\t")[0] for l in pu.sam_body(ref) if not l.startswith(b"@")))
    assert n_lines[""] == 56 and n_lines["-1"] == 29, n_lines


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_option_values_outside_their_domain(hostsim, workdir):
    """Where the reference dies of a signal (-A 0: assertion in ksw_align2; -c 0: division by zero; -w -1: abort) or runs
    on with values its loops were not written for, the engine returns status 1 with a message -- a signal would take the
    executor JVM down.  A negative -m reads as `no mate rescue` in the reference (bwa/bwamem_pair.c:267) and here."""
    c = pu.make_case(workdir, "pe_small")
    env = dict(os.environ, **pu.HOSTSIM_ENV)
    for bad in (["-A", "0"], ["-A", "-1"], ["-c", "0"], ["-c", "-1"], ["-w", "-1"], ["-k", "-3"], ["-B", "-3"], ["-O", "-2"], ["-E", "-1"]):
        r = subprocess.run([hostsim, "mem"] + bad + [c["ref"]] + c["fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert r.returncode == 1 and b"[E::main_mem]" in r.stderr, (bad, r.returncode)
        assert not [l for l in r.stdout.splitlines() if not l.startswith(b"@")], bad
    for opts in (["-A", "0"], ["-c", "0"], ["-w", "-1"]):   # what the reference does with them
        o = subprocess.run([pu.ORACLE, "mem"] + opts + [c["ref"]] + c["fq"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert o.returncode < 0 or o.returncode > 1, (opts, o.returncode)
    for same in (["-m", "-1"], ["-m", "-5"], ["-k", "0"], ["-E", "0"], ["-w", "0"], ["-B", "0"], ["-O", "0"], ["-c", "1"]):
        assert not pu.first_diff(pu.oracle_sam(c, extra=same), pu.engine_sam(c, hostsim, extra=same, env=pu.HOSTSIM_ENV)), same


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_input_read_errors_are_fatal_like_the_reference(hostsim, workdir):
    """err_gzread / err_gzclose (bwa/utils.c:214-225,284-289) end the reference's run with status 1: a .gz cut off in
    mid-stream is aligned as far as it decodes and then fails at gzclose (complete output, status 1); a directory given
    as a read file fails at the first gzread.  The engine used to report success on both."""
    import gzip
    c = pu.make_case(workdir, "pe_small")
    env = dict(os.environ, **pu.HOSTSIM_ENV)
    gz = os.path.join(c["dir"], "r_1_cut.fq.gz")
    with open(c["fq"][0], "rb") as fi, gzip.open(gz + ".full", "wb", compresslevel=1) as fo:
        fo.write(fi.read())
    with open(gz + ".full", "rb") as fi, open(gz, "wb") as fo:
        fo.write(fi.read()[:9000])
    o = subprocess.run([pu.ORACLE, "mem", c["ref"], gz], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert o.returncode == 1 and b"[gzclose]" in o.stderr
    out = os.path.join(c["dir"], "cut.sam")
    e = subprocess.run([hostsim, "mem", "-f", out, c["ref"], gz], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert e.returncode == 1 and b"[gzclose]" in e.stderr, e.stderr[-300:]
    ours = pu.sam_body(out)
    assert ours == [l for l in o.stdout.splitlines(True) if not l.startswith(b"@PG")] and len(ours) > 50
    e = subprocess.run([hostsim, "mem", "-f", out, c["ref"], c["dir"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert e.returncode == 1 and b"[gzread]" in e.stderr


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_paired_names_must_match(hostsim, workdir):
    """Mate files that are out of step: mem_sam_pe ends the reference's run on the first pair whose reads carry different
    names (bwa/bwamem_pair.c:355,385; it exits with status 1).  The engine returns 1 with the same message instead of
    aligning the wrong reads to each other, and leaves a well-formed (header-only) file.  Found by scripts/fuzz_parity.py."""
    c = pu.make_case(workdir, "pe_small")
    bad = os.path.join(c["dir"], "r_2_shifted.fq")
    recs = open(c["fq"][1], "rb").read().split(b"\n")
    with open(bad, "wb") as f:   # drop the 11th record of mate 2: every later pair is mismatched
        f.write(b"\n".join(recs[:40] + recs[44:]))
    r = subprocess.run([pu.ORACLE, "mem", c["ref"], c["fq"][0], bad], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"paired reads have different names" in r.stderr
    out = os.path.join(c["dir"], "shifted.sam")
    e = subprocess.run([hostsim, "mem", "-f", out, c["ref"], c["fq"][0], bad], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, **pu.HOSTSIM_ENV))
    assert e.returncode == 1
    assert b'paired reads have different names: "r10", "r11"' in e.stderr, e.stderr[-500:]
    assert all(l.startswith(b"@") for l in open(out, "rb").read().splitlines())


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_hostsim_smart_pairing(hostsim, workdir):
    """-p (bwa/fastmap.c:64-83): interleaved input with single-end reads mixed in, several -K batches."""
    c = pu.make_case(workdir, "pe_small")
    inter = pu.make_interleaved(c)
    c2 = dict(c, fq=[inter], opts=["-p", "-K", "30000"])
    ref = pu.oracle_sam(c2)
    our = pu.engine_sam(c2, hostsim, env=pu.HOSTSIM_ENV)
    assert not pu.first_diff(ref, our)
    assert len(pu.sam_body(our)) > 700


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_reader_paths_match_oracle(hostsim, workdir):
    """The reader takes a memory-mapped fast path for plain 4-line FASTQ and the kseq-exact general parser for
    everything else; all of them must cut the same batches and yield the same records as the reference's kseq:
    gzip input, the general parser forced on a plain file, CR/LF line ends, multi-line FASTA, and a file that
    switches between the two parsers record by record (FASTQ records followed by FASTA records)."""
    import gzip
    c = pu.make_case(workdir, "pe_small")
    d = c["dir"]
    ref = pu.oracle_sam(c)
    # (1) general parser forced on the plain files
    our = pu.engine_sam(c, hostsim, env=dict(pu.HOSTSIM_ENV, B200BWA_NO_MMAP="1"))
    assert not pu.first_diff(ref, our)
    # (2) gzip-compressed mates
    gz = []
    for f in c["fq"]:
        g = f + ".gz"
        with open(f, "rb") as fi, gzip.open(g, "wb", compresslevel=1) as fo:
            fo.write(fi.read())
        gz.append(g)
    our = pu.engine_sam(dict(c, fq=gz), hostsim, env=pu.HOSTSIM_ENV)
    assert not pu.first_diff(ref, our)
    # (3) CR/LF line ends, (4) multi-line FASTA after FASTQ records in one single-end file
    recs = open(c["fq"][0], "rb").read().split(b"\n")
    crlf = os.path.join(d, "crlf.fq")
    with open(crlf, "wb") as f:
        f.write(b"\r\n".join(recs))
    mixed = os.path.join(d, "mixed.fq")
    with open(mixed, "wb") as f:
        n_rec = (len(recs) - 1) // 4
        for k in range(n_rec):
            name, seq, _, qual = recs[4 * k:4 * k + 4]
            if k % 3 == 2:   # FASTA, sequence folded at 60 columns
                f.write(b">" + name[1:] + b" some comment\n")
                for p in range(0, len(seq), 60):
                    f.write(seq[p:p + 60] + b"\n")
            else:
                f.write(b"\n".join([name, seq, b"+" + (name[1:] if k % 5 == 0 else b""), qual]) + b"\n")
    for fq in (crlf, mixed):
        c1 = dict(c, fq=[fq], opts=["-K", "40000"])
        r1 = pu.oracle_sam(c1)
        o1 = pu.engine_sam(c1, hostsim, env=pu.HOSTSIM_ENV)
        assert not pu.first_diff(r1, o1), fq
        assert len(pu.sam_body(o1)) > 300


# ---- one job over several devices; parallel scan + device-side ingest -----------------------------
@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_hostsim_one_job_over_several_engines(hostsim, workdir):
    """b2_mem_main deals the -K batches of ONE job over G engines (batch b -> engine b mod G, each batch carrying its
    absolute n_processed) and writes one ordered SAM (kt_pipeline's ordering, bwa/kthread.c:92-102, bwa/fastmap.c:84-89).
    The host checker build pretends to have three devices; the SAM must not depend on G."""
    c = pu.make_case(workdir, "pe150")
    ref = pu.oracle_sam(c)
    for devs in ("0", "0,1,2", "1,1"):
        our = pu.engine_sam(c, hostsim, env=dict(pu.HOSTSIM_ENV, B200BWA_DEVICES=devs, B200BWA_WORKERS="2",
                                                 B200BWA_HOSTSIM_DEVICES="3"))
        assert not pu.first_diff(ref, our), devs
    # through a pipe (ordered-writer path instead of positioned writes)
    our = pu.engine_sam(c, hostsim, env=dict(pu.HOSTSIM_ENV, B200BWA_DEVICES="0,1,2", B200BWA_WORKERS="2"), via_f=False)
    assert not pu.first_diff(ref, our)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_parallel_scan_and_device_ingest(hostsim, workdir):
    """The mapped input is scanned in chunks by several threads, each guessing its first record start; the reader
    accepts a chunk only if it starts where the previous one ended.  Tiny chunks (every record straddles a boundary),
    one or many threads, SparkBWA's own file shape (a blank line after every record, BwaPairedAlignment.java:86-105) and
    a last record without newline must all give the batches -- and through the device-side packing (k_ingest_*: name
    trimming, nt4 codes) the SAM -- of the reference's kseq + bseq_read."""
    c = pu.make_case(workdir, "pe_small")
    sp = []
    for f in c["fq"]:
        g = f + ".spark"
        recs = open(f, "rb").read().split(b"\n")
        with open(g, "wb") as fo:
            n = (len(recs) - 1) // 4
            for k in range(n):
                fo.write(b"\n".join(recs[4 * k:4 * k + 4]) + b"\n" + (b"\n" if k < n - 1 else b""))
        sp.append(g)
    c_sp = dict(c, fq=sp)
    for K in ("30000", "7000"):
        ref = pu.oracle_sam(dict(c_sp, opts=["-K", K, "-C"]))
        assert not pu.first_diff(ref, pu.oracle_sam(dict(c, opts=["-K", K, "-C"])))
        for env in ({}, {"B200BWA_SCAN_CHUNK": "100"}, {"B200BWA_SCAN_CHUNK": "1000", "B200BWA_SCAN_THREADS": "7"},
                    {"B200BWA_SCAN_CHUNK": "64", "B200BWA_SCAN_THREADS": "1"}, {"B200BWA_NO_COMPACT": "1"}):
            for cc in (c, c_sp):
                r = subprocess.run([hostsim, "mem", "-K", K, "-C", "-f", os.path.join(c["dir"], "scan.sam"), cc["ref"]] + cc["fq"],
                                   env=dict(os.environ, B200BWA_TIMES="1", **pu.HOSTSIM_ENV, **env), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                assert r.returncode == 0, r.stderr[-1000:]
                assert not pu.first_diff(ref, os.path.join(c["dir"], "scan.sam")), (K, env)
                want = b"host pack" if "B200BWA_NO_COMPACT" in env else b"device ingest"
                assert want in r.stderr and (b"host pack" if want == b"device ingest" else b"device ingest") not in r.stderr


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_alt_contig_paths(hostsim, workdir):
    """ALT contigs (<prefix>.alt, bwa/bntseq.c:179-204): the fixture must actually reach the ALT-aware code -- @SQ ... AH:*
    header lines, pa:f: tags (bwa/bwamem.c:926: a primary that also has an ALT hit), ALT supplementary records
    (bwa/bwamem_pair.c:340-347) and XA lists beyond max_XA_hits (bwa/bwamem_extra.c:98-140, max_XA_hits_alt) -- and
    -j (ignore the .alt file) must give the plain result."""
    c = pu.make_case(workdir, "pe_alt")
    ref = pu.oracle_sam(c)
    body = pu.sam_body(ref)
    assert sum(1 for l in body if l.startswith(b"@SQ") and l.rstrip().endswith(b"AH:*")) == 8
    assert sum(1 for l in body if b"\tpa:f:" in l) > 50
    assert sum(1 for l in body if b"\tXA:Z:" in l) > 100
    assert not pu.first_diff(ref, pu.engine_sam(c, hostsim, env=pu.HOSTSIM_ENV))
    for extra in (["-j"], ["-h", "2,50"], ["-a", "-M"]):
        r = pu.oracle_sam(c, extra=extra)
        assert not pu.first_diff(r, pu.engine_sam(c, hostsim, extra=extra, env=pu.HOSTSIM_ENV)), extra


@pytest.mark.skipif(not os.access(pu.ORACLE_BATCH, os.X_OK), reason="oracle/_ref/oracle_batch not built")
def test_hostsim_absolute_read_index_and_int32_wrap(hostsim, workdir):
    """results depend on the absolute index of a batch's first read, incl. the int32 wrap of `id<<8`."""
    c = pu.make_case(workdir, "pe_repeat")
    assert pu.check_first_read_offsets(c, hostsim, env=pu.HOSTSIM_ENV) > 100


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_hostsim_config0_full_size(hostsim, workdir):
    """BASELINE.json configs[0] at its stated size (10 k x 100 bp single-end vs 5 Mbp), with the argv SparkBWA builds
    for a single-end partition (Bwa.java:289-390): bwa mem <-w tokens> -f <out> <index> <fq>."""
    c = pu.make_case(workdir, "cfg0_se100_5Mbp")
    ref = pu.oracle_sam(c)
    our = pu.engine_sam(c, hostsim, env=pu.HOSTSIM_ENV)
    assert not pu.first_diff(ref, our)
    assert len(pu.sam_body(our)) >= 10001


def test_reducer_semantics(tmp_path):
    """b200bwa_reduce_sam == the reducer of BwaInterpreter.java:342-376: files concatenated in order, lines starting with
    '@' kept from the first file only, every line newline-terminated, inputs deleted."""
    lib = ctypes.CDLL(LIB)
    files = []
    texts = [b"@SQ\tSN:c\tLN:9\n@PG\tID:bwa\nr1\t0\tc\t1\n@odd\t4\t*\n", b"@SQ\tSN:c\tLN:9\n@PG\tID:bwa\nr2\t0\tc\t2\nr3\t16\tc\t3",
             b"", b"@SQ\tSN:c\tLN:9\n"]
    for i, t in enumerate(texts):
        p = tmp_path / f"part-{i}.sam"
        p.write_bytes(t)
        files.append(str(p).encode())
    out = str(tmp_path / "FullOutput.sam")
    arr = (ctypes.c_char_p * len(files))(*files)
    assert lib.b200bwa_reduce_sam(len(files), arr, out.encode(), 1) == 0
    # Java twin: for i, file: for line in file: if i == 0 or not line.startswith("@"): out.write(line + "\n")
    want = b""
    for i, t in enumerate(texts):
        lines = t.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        want += b"".join(l + b"\n" for l in lines if i == 0 or not l.startswith(b"@"))
    assert open(out, "rb").read() == want
    assert not any(os.path.exists(f.decode()) for f in files)
    assert lib.b200bwa_reduce_sam(1, (ctypes.c_char_p * 1)(b"/nonexistent.sam"), out.encode(), 0) == 1


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_index_builder_matches_bwa_index(hostsim, workdir):
    """the index builder (b2_index.cu: suffix groups, radix sort, tie refinement; b2_host.cpp: FASTA
    packing with the lrand48 replay for ambiguous bases, .ann/.amb) against the reference's `bwa index`, byte for byte,
    with the kernel bodies run by the host checker build."""
    for tag, ctg, with_n, env in pu.index_cases():
        pu.check_index_builder(hostsim, os.path.join(workdir, "index"), tag, ctg, with_n, env)


@pytest.mark.skipif(not pu.have_oracle(), reason="oracle/_ref not built")
def test_hostsim_concurrent_calls_lease_engines(hostsim, workdir):
    """Several task threads of one executor calling the entry point at once (the reference cannot: global state): each
    call leases the free engines of its device list -- one each with B200BWA_DEVICES_PER_CALL=1, turns otherwise -- and
    every call writes the reference's SAM."""
    c = pu.make_case(workdir, "pe150")
    ref = pu.oracle_sam(c)
    out = os.path.join(c["dir"], "conc.sam")
    for env in ({"B200BWA_DEVICES": "0,1,2", "B200BWA_DEVICES_PER_CALL": "1"}, {"B200BWA_DEVICES": "0,1"}, {"B200BWA_DEVICES": "0"}):
        e = dict(os.environ, **pu.HOSTSIM_ENV, B200BWA_HOSTSIM_DEVICES="3", B200BWA_WORKERS="2", B200BWA_CLI_CONCURRENT="3", **env)
        r = subprocess.run([hostsim, "mem"] + c["opts"] + ["-f", out, c["ref"]] + c["fq"], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr[-1000:]
        for k in range(3):
            assert not pu.first_diff(ref, out + f".{k}"), (env, k)
