"""world_size-2 tests of the multi-GPU path on CPU (gloo): batch sharding + ordered merge reproduce the
single-rank SAM, and the index image broadcast delivers identical bytes to every rank."""
import os
import subprocess
import sys

import pytest

import parity_util as pu

WORKER = r'''
import os, sys, subprocess
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import parity_util as pu
from sparkbwa_b200 import shard
work, binary = sys.argv[2], sys.argv[3]
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
gold = os.path.join(pu.ROOT, "tests", "golden", "tiny")
# (1) index broadcast: rank 0 reads the files, everyone ends up with the same image
imgs = []
for ext in ("bwt", "sa", "pac"):
    a = np.fromfile(os.path.join(gold, "ref.fa." + ext), dtype=np.uint8)
    t = torch.from_numpy(a.copy()) if rank == 0 else torch.zeros(a.size, dtype=torch.uint8)
    dist.broadcast(t, src=0)
    assert bytes(t.numpy().tobytes()) == a.tobytes()
# (2) every rank aligns its share of the -K batches
out = os.path.join(work, f"piece{rank}.sam")
env = dict(os.environ, **pu.HOSTSIM_ENV, **shard.shard_env(rank, world))
r = subprocess.run([binary, "mem", "-K", "20000", "-f", out, os.path.join(gold, "ref.fa"), os.path.join(gold, "pe_1.fq"),
                    os.path.join(gold, "pe_2.fq")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
assert r.returncode == 0, r.stderr.decode()[-2000:]
dist.barrier()
if rank == 0:
    n = shard.merge_shards([os.path.join(work, f"piece{k}.sam") for k in range(world)], os.path.join(work, "merged.sam"))
    assert n >= 3, n
    assert pu.sam_body(os.path.join(work, "merged.sam")) == pu.sam_body(os.path.join(gold, "pe.sam"))
    print("MERGED_OK", n)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_matches_single(tmp_path):
    subprocess.run(["make", "-C", os.path.join(pu.ROOT, "tests", "hostsim"), "-j4"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.PIPE)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", str(script), pu.ROOT, str(tmp_path), pu.HOSTSIM],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, (r.stdout.decode()[-1500:], r.stderr.decode()[-3000:])
    assert b"MERGED_OK" in r.stdout


def test_shard_assignment_is_a_partition():
    from sparkbwa_b200 import shard
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in shard.my_batches(37, r, world))
        assert seen == list(range(37))
