"""Seeded synthetic references and wgsim-style reads.

The reference ships no data, tests or read simulator, so every workload of BASELINE.json is
regenerated from a seed: an i.i.d. uniform ACGT reference with planted dispersed repeats (so that
`max_occ`, B-tree splits and mate rescue are exercised) and no `N` (keeps `.pac` independent of
`lrand48`, bntseq.c:290), and reads with substitutions, geometric indels, constant quality,
FR paired ends with N(mean, sd) outer distance and `/1` `/2` name suffixes (trimmed by the
reader exactly like bwa.c:27-31).
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_reference(total_len: int, n_contigs: int = 1, seed: int = 1, repeat_frac: float = 0.02,
                   rep_len=(300, 5000), rep_copies=(2, 50), rep_div=0.03):
    """Return a list of (name, codes uint8[0..3]) contigs."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=total_len, dtype=np.uint8)
    planted = 0
    budget = int(total_len * repeat_frac)
    while planted < budget and total_len > 4 * rep_len[0]:
        L = int(rng.integers(rep_len[0], min(rep_len[1], max(rep_len[0] + 1, total_len // 8)) + 1))
        ncopy = int(rng.integers(rep_copies[0], rep_copies[1] + 1))
        src = int(rng.integers(0, total_len - L))
        fam = codes[src:src + L].copy()
        div = float(rng.uniform(0, rep_div))
        for _ in range(ncopy):
            cp = fam.copy()
            nmut = rng.binomial(L, div)
            if nmut:
                pos = rng.integers(0, L, size=nmut)
                cp[pos] = (cp[pos] + rng.integers(1, 4, size=nmut, dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                cp = (3 - cp[::-1]).astype(np.uint8)
            dst = int(rng.integers(0, total_len - L))
            codes[dst:dst + L] = cp
            planted += L
            if planted >= budget:
                break
    # cut into contigs of slightly unequal length
    bounds = [0]
    base = total_len // n_contigs
    for i in range(1, n_contigs):
        bounds.append(i * base + int(rng.integers(-base // 8, base // 8 + 1)))
    bounds.append(total_len)
    return [(f"chr{i + 1}", codes[bounds[i]:bounds[i + 1]]) for i in range(n_contigs)]


def add_alt_contigs(contigs, n_alt: int = 4, alt_len=(1500, 4000), div: float = 0.02, seed: int = 5):
    """ALT haplotypes: near-copies (substitutions at rate `div` plus one small indel) of segments of the primary
    contigs, appended as contigs named <chr>_alt<k>.  Returns (all contigs, [ALT names]) -- the names go into
    <prefix>.alt (bwa/bntseq.c:179-204) so that the ALT-aware branches run (bwa/bwamem.c:535-551,
    bwa/bwamem_pair.c:340-347, bwa/bwamem_extra.c:98-140)."""
    rng = np.random.default_rng(seed)
    out = list(contigs)
    names = []
    for k in range(n_alt):
        ci = int(rng.integers(0, len(contigs)))
        name, codes = contigs[ci]
        L = int(rng.integers(alt_len[0], alt_len[1] + 1))
        L = min(L, len(codes) // 2)
        st = int(rng.integers(0, len(codes) - L))
        cp = codes[st:st + L].copy()
        nmut = rng.binomial(L, div)
        if nmut:
            pos = rng.integers(0, L, size=nmut)
            cp[pos] = (cp[pos] + rng.integers(1, 4, size=nmut, dtype=np.uint8)) & 3
        cut = int(rng.integers(L // 4, 3 * L // 4))
        if k % 2 == 0:   # a short deletion
            cp = np.concatenate([cp[:cut], cp[cut + int(rng.integers(1, 9)):]])
        else:            # a short insertion
            cp = np.concatenate([cp[:cut], rng.integers(0, 4, size=int(rng.integers(1, 9)), dtype=np.uint8), cp[cut:]])
        if k % 3 == 2:
            cp = (3 - cp[::-1]).astype(np.uint8)
        nm = f"{name}_alt{k}"
        out.append((nm, cp))
        names.append(nm)
    return out, names


def write_fasta(path: str, contigs, width: int = 60):
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(b">" + name.encode() + b"\n")
            s = _BASES[codes]
            n = len(s)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def _mutate(rng, frag, sub, indel, ext=0.3):
    """Apply substitutions and geometric indels; returns codes (may change length)."""
    L = len(frag)
    out = frag.copy()
    if sub > 0:
        m = rng.random(L) < sub
        k = int(m.sum())
        if k:
            out[m] = (out[m] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    if indel > 0:
        n_ev = rng.binomial(L, indel)
        if n_ev:
            pieces = []
            pos = np.sort(rng.integers(1, L - 1, size=n_ev))
            last = 0
            for p in pos:
                if p < last:
                    continue
                glen = int(rng.geometric(1 - ext))
                pieces.append(out[last:p])
                if rng.random() < 0.5:  # insertion
                    pieces.append(rng.integers(0, 4, size=glen, dtype=np.uint8))
                    last = p
                else:  # deletion
                    last = min(L, p + glen)
            pieces.append(out[last:])
            out = np.concatenate(pieces)
    return out


def simulate_reads(contigs, n: int, read_len: int, paired: bool, seed: int = 7, sub: float = 0.01,
                   indel: float = 0.001, ins_mean: float = 400.0, ins_sd: float = 40.0,
                   n_frac: float = 0.0, mate2_sub: float | None = None, junk_frac: float = 0.0,
                   chimeric_frac: float = 0.0, ragged: bool = False):
    """Return (reads1, reads2|None); each a list of (name bytes, seq bytes, qual bytes)."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c[1]) for c in contigs], dtype=np.int64)
    prob = lens / lens.sum()
    r1, r2 = [], []
    slack = 64
    for i in range(n):
        ci = int(rng.choice(len(contigs), p=prob))
        ref = contigs[ci][1]
        if paired:
            isz = int(max(read_len + 10, rng.normal(ins_mean, ins_sd)))
        else:
            isz = read_len
        isz = min(isz, len(ref) - slack - 1)
        st = int(rng.integers(0, len(ref) - isz - slack))
        frag = ref[st:st + isz + slack]
        flip = rng.random() < 0.5
        name = f"r{i}".encode()
        if rng.random() < junk_frac:
            a = rng.integers(0, 4, size=read_len, dtype=np.uint8)
            b = rng.integers(0, 4, size=read_len, dtype=np.uint8)
        else:
            left = _mutate(rng, frag[:read_len + slack // 2], sub, indel)[:read_len]
            if paired:
                rgt_src = (3 - frag[:isz][::-1]).astype(np.uint8)
                right = _mutate(rng, rgt_src[:read_len + slack // 2] if len(rgt_src) >= read_len + slack // 2
                                else rgt_src, sub if mate2_sub is None else mate2_sub, indel)[:read_len]
            else:
                right = None
            if flip:
                if paired:
                    a, b = right, left
                else:
                    a, b = (3 - left[::-1]).astype(np.uint8), None
            else:
                a, b = left, right
        if chimeric_frac > 0 and rng.random() < chimeric_frac:  # splice in a segment from elsewhere
            cj = int(rng.choice(len(contigs), p=prob))
            other = contigs[cj][1]
            cut = int(rng.integers(30, read_len - 30))
            st2 = int(rng.integers(0, len(other) - read_len))
            seg = other[st2:st2 + read_len - cut]
            if rng.random() < 0.5:
                seg = (3 - seg[::-1]).astype(np.uint8)
            a = np.concatenate([a[:cut], seg])
        if ragged:
            a = a[:int(rng.integers(1, read_len + 1))]
        for seqc, dst, sfx in ((a, r1, b"/1"), (b, r2, b"/2")):
            if seqc is None:
                continue
            s = _BASES[seqc].copy()
            if n_frac > 0:
                m = rng.random(len(s)) < n_frac
                s[m] = ord("N")
            dst.append((name + (sfx if paired else b""), s.tobytes(), b"I" * len(s)))
    return r1, (r2 if paired else None)


def write_fastq(path: str, reads):
    with open(path, "wb") as f:
        for name, seq, qual in reads:
            f.write(b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n")


def simulate_pairs_fast(contigs, n_pairs: int, read_len: int = 150, seed: int = 7, sub: float = 0.01,
                        indel: float = 0.001, ins_mean: float = 400.0, ins_sd: float = 40.0, chunk: int = 200_000,
                        indel_rounds: int = 1):
    """Vectorised paired-end simulator for the benchmark workloads (millions of pairs).

    Same model as simulate_reads (uniform start, 50/50 strand, per-base substitutions, FR pairs with
    N(ins_mean, ins_sd) outer distance) with at most `indel_rounds` single-base indels per read (each round plants one
    with probability indel * read_len / indel_rounds; 1 for the 2x150 bp workloads, more for the 250 bp / 5 % one).  Yields chunks
    (seq1, seq2) of ASCII uint8 arrays shaped (m, read_len); names are r%09d by global pair index.
    """
    rng = np.random.default_rng(seed)
    ref = np.concatenate([c for _, c in contigs]).astype(np.uint8)
    lens = np.array([len(c) for _, c in contigs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    prob = lens / lens.sum()
    L = read_len
    cols = np.arange(L, dtype=np.int64)
    done = 0
    while done < n_pairs:
        m = min(chunk, n_pairs - done)
        ci = rng.choice(len(contigs), size=m, p=prob)
        isz = np.maximum(L + 10, rng.normal(ins_mean, ins_sd, size=m)).astype(np.int64)
        isz = np.minimum(isz, lens[ci] - 4)
        st = offs[ci] + (rng.random(m) * (lens[ci] - isz - 2)).astype(np.int64)
        out = []
        for which in range(2):
            if which == 0:
                idx = st[:, None] + cols[None, :]
            else:
                idx = (st + isz - 1)[:, None] - cols[None, :]
            # optional single-base indels: shift the source index after position p
            planted = []
            for _ in range(indel_rounds):
                ev = rng.random(m) < indel * L / indel_rounds
                p = rng.integers(10, L - 10, size=m)
                is_del = rng.random(m) < 0.5
                shift = (cols[None, :] >= p[:, None]) & ev[:, None]
                step = np.where(is_del, 1, -1)[:, None] * (1 if which == 0 else -1)
                idx = idx + shift * step
                planted.append((ev & ~is_del, p))
            idx = np.clip(idx, 0, len(ref) - 1)
            r = ref[idx]
            if which == 1:
                r = 3 - r
            for ins, p in planted:
                if ins.any():  # the inserted base itself is random
                    rows = np.nonzero(ins)[0]
                    r[rows, p[rows]] = rng.integers(0, 4, size=rows.size, dtype=np.uint8)
            if sub > 0:
                msk = rng.random((m, L)) < sub
                k = int(msk.sum())
                r[msk] = (r[msk] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
            out.append(r.astype(np.uint8))
        flip = rng.random(m) < 0.5
        a = np.where(flip[:, None], out[1], out[0])
        b = np.where(flip[:, None], out[0], out[1])
        yield _BASES[a], _BASES[b]
        done += m


def write_fastq_fixed(path: str, seq_chunks, first_index: int = 0, suffix: bytes = b"/1"):
    """Write fixed-width FASTQ records from (m, L) ASCII arrays (to a path, or appended to an open binary file);
    returns the number of reads."""
    import contextlib
    n = 0
    with (contextlib.nullcontext(path) if hasattr(path, "write") else open(path, "wb")) as f:
        for s in seq_chunks:
            m, L = s.shape
            ids = np.arange(first_index + n, first_index + n + m)
            name = np.char.zfill(ids.astype(str), 9).astype("S9").view(np.uint8).reshape(m, 9)
            head = np.empty((m, 1 + 1 + 9 + len(suffix) + 1), dtype=np.uint8)
            head[:, 0] = ord("@"); head[:, 1] = ord("r"); head[:, 2:11] = name
            head[:, 11:11 + len(suffix)] = np.frombuffer(suffix, dtype=np.uint8)
            head[:, -1] = 10
            mid = np.frombuffer(b"\n+\n", dtype=np.uint8)
            rec = np.concatenate([head, s, np.broadcast_to(mid, (m, 3)), np.full((m, L), ord("I"), dtype=np.uint8),
                                  np.full((m, 1), 10, dtype=np.uint8)], axis=1)
            f.write(rec.tobytes())
            n += m
    return n
