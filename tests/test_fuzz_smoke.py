"""A short slice of the differential fuzz campaigns (scripts/fuzz_parity.py, scripts/fuzz_index.py) inside the CPU suite:
random references, reads, `bwa mem` option strings and file formats, the engine's kernel bodies (host build) against
the unmodified reference.  The campaigns proper run thousands of trials."""
import os
import sys

import pytest

import parity_util as pu

sys.path.insert(0, os.path.join(pu.ROOT, "scripts"))

needs = pytest.mark.skipif(not (pu.have_oracle() and os.access(pu.HOSTSIM, os.X_OK)), reason="oracle/_ref or tests/hostsim not built")


@needs
@pytest.mark.parametrize("first,profile", [(1000, "regular"), (1000100, "corner-heavy"), (3000000, "file formats"), (3600000, "file formats")])
def test_mem_fuzz_slice(tmp_path, first, profile):
    import fuzz_parity as fz
    for seed in range(first, first + 8):
        rec = fz.trial((seed, str(tmp_path), False, "hostsim"))
        assert rec["ok"], (profile, rec)


@needs
def test_index_fuzz_slice(tmp_path):
    import fuzz_index as fi
    for seed in range(40):
        _, ok, msg = fi.trial((seed, str(tmp_path), "hostsim"))
        assert ok, (seed, msg)
