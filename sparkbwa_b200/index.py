"""FM-index construction: emits the same files `bwa index` writes.

Not part of the alignment hot path: the suffix array is obtained by prefix doubling with
`torch.sort` (library plumbing; runs on the GPU when one is present, on the CPU for small test
genomes) and everything else is element-wise tensor work.  The BWT/Occ/SA layout restates
bwa/bwtindex.c:65-124 (pac -> BWT over forward + reverse-complement text), :151-173
(bwt_bwtupdate_core: cumulative A/C/G/T counts interleaved every 128 symbols), bwa/bwt.c:62-84
(bwt_cal_sa, one sample every 32 rows), bwa/bwt.c:385-405 (file headers) and bwa/bntseq.c:66-96,
:275-328 (.ann/.amb/.pac).  BWT and SA are uniquely determined by the text, so the files are
byte-identical to the reference's for references without N (N needs an lrand48 replay,
bntseq.c:290 -- rejected here).
"""
from __future__ import annotations

import os

import numpy as np
import torch


def suffix_array(text: torch.Tensor) -> torch.Tensor:
    """SA of text+'$' (codes 0..3; '$' smallest).  Returns int64[n+1] with SA[0] = n."""
    dev = text.device
    n = int(text.numel())
    m = n + 1
    t = torch.zeros(m, dtype=torch.int64, device=dev)
    t[:n] = text.to(torch.int64) + 1
    K = 26  # 5**26 < 2**63
    key = torch.zeros(m, dtype=torch.int64, device=dev)
    for j in range(K):
        key.mul_(5)
        if j < m:
            key[: m - j].add_(t[j:])
    del t

    def ranks_of(keys):
        vals, order = torch.sort(keys)
        new = torch.ones(m, dtype=torch.int64, device=dev)
        new[0] = 0
        new[1:] = (vals[1:] != vals[:-1]).to(torch.int64)
        dense = torch.cumsum(new, 0)
        rank = torch.empty(m, dtype=torch.int64, device=dev)
        rank[order] = dense
        return rank, order, int(dense[-1].item())

    rank, order, mx = ranks_of(key)
    del key
    h = K
    while mx < m - 1:
        nxt = torch.zeros(m, dtype=torch.int64, device=dev)
        if h < m:
            nxt[: m - h] = rank[h:] + 1
        key = rank * (m + 1) + nxt
        del nxt
        rank, order, mx = ranks_of(key)
        del key
        h *= 2
    return order


def build_index(contigs, prefix: str, device: str | None = None, sa_intv: int = 32) -> dict:
    """contigs: list of (name, uint8 codes 0..3).  Writes prefix.{pac,ann,amb,bwt,sa}."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    fwd = np.concatenate([c for _, c in contigs]).astype(np.uint8)
    if fwd.max(initial=0) > 3:
        raise ValueError("references with N are not supported by this builder")
    l_pac = int(fwd.size)
    # ---- .pac (forward only) / .ann / .amb ------------------------------------------------------
    pad = (-l_pac) % 4
    f4 = np.concatenate([fwd, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    pac = (f4[:, 0] << 6 | f4[:, 1] << 4 | f4[:, 2] << 2 | f4[:, 3]).astype(np.uint8)
    with open(prefix + ".pac", "wb") as f:
        f.write(pac.tobytes())
        if l_pac % 4 == 0:
            f.write(b"\0")
        f.write(bytes([l_pac % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write(f"{l_pac} {len(contigs)} 11\n")
        off = 0
        for name, c in contigs:
            f.write(f"0 {name} (null)\n{off} {len(c)} 0\n")
            off += len(c)
    with open(prefix + ".amb", "w") as f:
        f.write(f"{l_pac} {len(contigs)} 0\n")
    # ---- BWT over forward + reverse complement ---------------------------------------------------
    text = torch.from_numpy(np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])).to(device)
    n = int(text.numel())
    sa = suffix_array(text)  # rows 0..n
    primary = int(torch.nonzero(sa == 0)[0].item())
    prev = torch.where(sa > 0, sa - 1, torch.zeros_like(sa))
    bwt_rows = text[prev]  # symbol preceding each suffix ('$' row gets a dummy)
    keep = torch.ones(n + 1, dtype=torch.bool, device=device)
    keep[primary] = False
    sym = bwt_rows[keep].to(torch.int64)  # n symbols, '$' removed
    del bwt_rows, prev, keep
    counts = torch.bincount(text.to(torch.int64), minlength=4).cpu().numpy()
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = np.cumsum(counts).astype(np.uint64)
    # pack 16 symbols per word, most significant pair first
    n_words = (n + 15) >> 4
    padn = n_words * 16 - n
    s16 = torch.cat([sym, torch.zeros(padn, dtype=torch.int64, device=device)]).reshape(-1, 16)
    shifts = torch.arange(15, -1, -1, device=device, dtype=torch.int64) * 2
    words = (s16 << shifts).sum(1)  # int64 holding a u32
    del s16
    # cumulative counts before every 128-symbol block, plus the closing totals
    n_blocks = (n + 127) // 128
    cnt = torch.zeros((n_blocks + 1, 4), dtype=torch.int64, device=device)
    for c in range(4):
        cs = torch.cumsum((sym == c).to(torch.int64), 0)
        idx = torch.arange(1, n_blocks + 1, device=device, dtype=torch.int64) * 128
        idx.clamp_(max=n)
        cnt[1:, c] = cs[idx - 1]
        del cs
    del sym
    wpad = n_blocks * 8 - n_words
    w8 = torch.cat([words, torch.zeros(wpad, dtype=torch.int64, device=device)]).reshape(n_blocks, 8)
    blk = torch.zeros((n_blocks, 16), dtype=torch.int64, device=device)
    blk[:, 0:8:2] = cnt[:n_blocks] & 0xFFFFFFFF
    blk[:, 1:8:2] = cnt[:n_blocks] >> 32
    blk[:, 8:] = w8
    flat = blk.reshape(-1)[: n_blocks * 16 - wpad]
    tail = torch.zeros(8, dtype=torch.int64, device=device)
    tail[0::2] = cnt[n_blocks] & 0xFFFFFFFF
    tail[1::2] = cnt[n_blocks] >> 32
    data = torch.cat([flat, tail]).cpu().numpy().astype(np.uint32)
    with open(prefix + ".bwt", "wb") as f:
        f.write(np.array([primary], dtype=np.uint64).tobytes())
        f.write(L2[1:].tobytes())
        f.write(data.tobytes())
    # ---- sampled SA --------------------------------------------------------------------------------
    samp = sa[::sa_intv].cpu().numpy().astype(np.uint64)
    n_sa = (n + sa_intv) // sa_intv
    assert samp.size == n_sa
    with open(prefix + ".sa", "wb") as f:
        f.write(np.array([primary], dtype=np.uint64).tobytes())
        f.write(L2[1:].tobytes())
        f.write(np.array([sa_intv, n], dtype=np.uint64).tobytes())
        f.write(samp[1:].tobytes())
    return {"l_pac": l_pac, "seq_len": n, "primary": primary, "n_sa": n_sa}
