"""Read-batch sharding across GPUs and the ordered merge of the per-rank SAM pieces.

The alignment path shards by `-K` batch: batch b goes to rank b % world, carries its absolute
`n_processed`, and is aligned independently (its own mem_pestat).  Each rank runs the unchanged entry
point with B200BWA_SHARD=rank/world and writes `<out>` plus `<out>.batches`; merging the pieces in
batch order reproduces the single-GPU (and the reference's) output byte for byte, the way
kt_pipeline's ordered third step does (bwa/kthread.c:92-102).
"""
from __future__ import annotations

import os


def shard_env(rank: int, world: int, device: int | None = None) -> dict:
    env = {"B200BWA_SHARD": f"{rank}/{world}"}
    if device is not None:
        env["B200BWA_DEVICE"] = str(device)
    return env


def my_batches(n_batches: int, rank: int, world: int):
    return [b for b in range(n_batches) if b % world == rank]


def merge_shards(paths, out_path: str) -> int:
    """paths[r] = SAM piece of rank r (rank 0 holds the header).  Returns the number of batches."""
    pieces = {}
    header = b""
    for r, p in enumerate(paths):
        with open(p + ".batches") as f:
            rows = [tuple(int(x) for x in line.split()) for line in f if line.strip()]
        with open(p, "rb") as f:
            data = f.read()
        body_len = sum(x[2] for x in rows)
        hdr_len = len(data) - body_len
        if r == 0:
            header = data[:hdr_len]
        elif hdr_len != 0:
            raise ValueError(f"{p}: unexpected header in a non-zero rank piece")
        off = hdr_len
        for idx, _n, ln in rows:
            if idx in pieces:
                raise ValueError(f"batch {idx} present twice")
            pieces[idx] = data[off:off + ln]
            off += ln
    n = len(pieces)
    if sorted(pieces) != list(range(n)):
        raise ValueError("missing batches: have %s" % sorted(pieces))
    with open(out_path, "wb") as f:
        f.write(header)
        for b in range(n):
            f.write(pieces[b])
    return n
