"""Manual driver: python tests/run_parity.py {hostsim|gpu} [case ...] -- full-SAM diff vs the oracle."""
import sys, tempfile, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_util as pu

mode = sys.argv[1]
cases = sys.argv[2:] or list(pu.CASES)
work = os.environ.get("B2_WORK", os.path.join(tempfile.gettempdir(), "b2cases"))
bad = 0
for name in cases:
    c = pu.make_case(work, name)
    t = time.time(); ref = pu.oracle_sam(c); t_or = time.time() - t
    t = time.time()
    try:
        if mode == "hostsim":
            our = pu.engine_sam(c, pu.HOSTSIM, env=pu.HOSTSIM_ENV)
        else:
            our = pu.engine_sam(c, pu.CLI)
    except Exception as e:
        print(name, "ENGINE FAILED", e); bad += 1; continue
    t_en = time.time() - t
    d = pu.first_diff(ref, our)
    n = len(pu.sam_body(ref))
    print(f"{name}: lines={n} oracle={t_or:.2f}s engine={t_en:.2f}s {'IDENTICAL' if not d else 'DIFF'}")
    for i, x, y in d:
        print("  line", i); print("   ref:", x.rstrip()); print("   our:", y.rstrip())
    bad += bool(d)
sys.exit(1 if bad else 0)
