#!/usr/bin/env python
"""Test infrastructure (CPU): the host build of the kernel bodies and the host pipeline under AddressSanitizer +
UndefinedBehaviorSanitizer and under ThreadSanitizer, on the parity cases of tests/parity_util.py.

  ASan/UBSan  every case with ordinary capacities and with deliberately tiny ones (every overflow / retry path of the
              stages; the class of fault that round 2's 8-GPU run hit on the device), SAM compared with the oracle;
  TSan        the host pipeline: one job over three pretended devices x three workers, three concurrent calls leasing
              engines, the ordered pipe writer, a failing run, smart pairing, gzip input, tiny capacities, the index builder.

Needs a g++ with the sanitizer runtimes (the system /usr/bin/g++-13 here; /opt/gcc has none).
  python scripts/sanitize_hostsim.py [--out /tmp/b2san] [case ...]
"""
import argparse
import gzip
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402

SRC = os.path.join(ROOT, "sparkbwa_b200", "csrc")
UNITS = [("engine", "b2_engine.cu"), ("index", "b2_index.cu"), ("b2_host", "b2_host.cpp"), ("b2_capi", "b2_capi.cpp"), ("b2_cli", "b2_cli.cpp")]


def build(cxx, out, san):
    os.makedirs(out, exist_ok=True)
    base = ["-O1", "-g", "-std=c++17", "-DB2_HOSTSIM", "-ffp-contract=off", "-fno-omit-frame-pointer", "-Wno-unused-function",
            "-Wno-sign-compare", "-I" + SRC, "-I" + os.path.join(ROOT, "tests", "hostsim")]
    with ThreadPoolExecutor(5) as ex:
        list(ex.map(lambda u: subprocess.run([cxx] + base + san + ["-x", "c++", "-c", os.path.join(SRC, u[1]), "-o", os.path.join(out, u[0] + ".o")],
                                             check=True), UNITS))
    exe = os.path.join(out, "b200bwa_hostsim")
    subprocess.run([cxx] + san + ["-o", exe] + [os.path.join(out, u[0] + ".o") for u in UNITS] + ["-lz", "-lpthread"], check=True)
    return exe


def run(exe, args, env, marker):
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    hits = [l for l in r.stderr.decode(errors="replace").splitlines() if any(m in l for m in marker)]
    return r, hits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/b2san")
    ap.add_argument("--cxx", default="/usr/bin/g++-13")
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    asan = build(a.cxx, os.path.join(a.out, "asan"), ["-fsanitize=address,undefined"])
    tsan = build(a.cxx, os.path.join(a.out, "tsan"), ["-fsanitize=thread"])
    work = os.path.join(a.out, "cases")
    bad = 0
    tiny = [{}, dict(B200BWA_SCRATCH="512", B200BWA_CAP_INTV="2", B200BWA_SLOT="64"),
            dict(B200BWA_SCRATCH="2048", B200BWA_BIG_SLOT="4096", B200BWA_BIG_POOL=str(64 << 10)),
            dict(B200BWA_SCRATCH="1024", B200BWA_BIG_POOL="0")]
    for name in (a.cases or list(pu.CASES)):
        c = pu.make_case(work, name)
        ref = pu.oracle_sam(c)
        out = os.path.join(c["dir"], "san.sam")
        for k, extra in enumerate(tiny):
            env = dict(os.environ, **pu.HOSTSIM_ENV, ASAN_OPTIONS="detect_leaks=0:halt_on_error=0")
            env.update(extra)
            r, hits = run(asan, ["mem"] + c["opts"] + ["-f", out, c["ref"]] + c["fq"], env, ("AddressSanitizer", "runtime error"))
            same = r.returncode == 0 and not pu.first_diff(ref, out)
            print(f"asan {name} capacities#{k}: rc {r.returncode} identical {same} reports {len(hits)}", flush=True)
            bad += (not same) + len(hits)
    # the host pipeline under TSan
    c = pu.make_case(work, "pe_small")
    inter = pu.make_interleaved(c)
    gz = []
    for f in c["fq"]:
        with open(f, "rb") as fi, gzip.open(f + ".gz", "wb", compresslevel=1) as fo:
            fo.write(fi.read())
        gz.append(f + ".gz")
    shifted = os.path.join(c["dir"], "r_2_shifted.fq")
    recs = open(c["fq"][1], "rb").read().split(b"\n")
    open(shifted, "wb").write(b"\n".join(recs[:1200] + recs[1204:]))
    base = dict(os.environ, **pu.HOSTSIM_ENV, TSAN_OPTIONS="halt_on_error=0", B200BWA_DEVICES="0,1,2", B200BWA_WORKERS="3", B200BWA_HOSTSIM_DEVICES="3")
    out = os.path.join(c["dir"], "tsan.sam")
    scen = [("one job, 3 devices x 3 workers", ["mem", "-K", "20000", "-f", out, c["ref"]] + c["fq"], {}),
            ("3 concurrent calls", ["mem", "-K", "20000", "-f", out, c["ref"]] + c["fq"], dict(B200BWA_CLI_CONCURRENT="3", B200BWA_DEVICES_PER_CALL="1")),
            ("ordered pipe writer", ["mem", "-K", "20000", c["ref"]] + c["fq"], {}),
            ("failing run (names differ)", ["mem", "-K", "20000", "-f", out, c["ref"], c["fq"][0], shifted], {}),
            ("smart pairing", ["mem", "-p", "-K", "20000", "-f", out, c["ref"], inter], {}),
            ("gzip input", ["mem", "-K", "20000", "-f", out, c["ref"]] + gz, {}),
            ("tiny capacities", ["mem", "-K", "20000", "-f", out, c["ref"]] + c["fq"], dict(B200BWA_SCRATCH="1024", B200BWA_BIG_SLOT="4096", B200BWA_BIG_POOL="65536", B200BWA_CAP_INTV="2", B200BWA_SLOT="64")),
            ("index builder", ["index", "-p", os.path.join(c["dir"], "tsan_idx"), c["ref"]], {})]
    for what, args, extra in scen:
        env = dict(base)
        env.update(extra)
        r, hits = run(tsan, args, env, ("WARNING: ThreadSanitizer",))
        print(f"tsan {what}: rc {r.returncode} reports {len(hits)}", flush=True)
        bad += len(hits)
    print("sanitizers:", "clean" if not bad else f"{bad} findings")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
