// FM-index primitives and SMEM seeding over the HBM-resident BWT+Occ image.
//
// Semantics restated from the reference (not its code): rank queries bwa/bwt.c:107-129 (bwt_occ),
// :169-186 (bwt_occ4), :189-220 (bwt_2occ4); bidirectional extension :262-275 (bwt_extend);
// SMEM search :289-356 (bwt_smem1a with max_intv == 0, which is how bwamem.c:121,137 call it);
// third-pass seeds :358-379 (bwt_seed_strategy1); three-pass collection bwamem.c:114-162
// (mem_collect_intv); SA lookup bwt.c:53-59,86-96 (bwt_invPsi, bwt_sa).
//
// Layout (verbatim .bwt payload): block b = 64 bytes at bwt + 16*b words:
//   bytes  0..31 : cumulative counts of A,C,G,T before the block (4 x u64)
//   bytes 32..63 : 128 symbols, 2 bits each, 16 per u32, most-significant pair first.
// A rank query at k counts symbols [0, (k&127)] of block k>>7 (k already shifted past '$').
//
// On the device a "seeding group" of B2_SEED_LANES lanes serves one read: every lane runs the same
// control flow and lane g fetches 16-byte chunk g of the two blocks a bidirectional extension
// touches (one LDG.128 per lane per block => one 64-byte request per block per group), popcounts
// its part and the partial ranks are exchanged with warp shuffles.
#pragma once
#include "b2_types.h"

#define B2_SEED_LANES 4

B2_HD inline uint32_t b2_cnt_word(uint32_t w, int m) {
  // packed counts (A | C<<8 | G<<16 | T<<24) of the top m (0..16) symbols of a BWT word
  if (m <= 0) return 0;
  uint32_t vm = 0x55555555u & (m >= 16 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (32 - 2 * m)));
  uint32_t hi = (w >> 1), lo = w;
#if defined(__CUDA_ARCH__)
  uint32_t t = __popc(hi & lo & vm), g = __popc(hi & ~lo & vm), c = __popc(~hi & lo & vm);
#else
  uint32_t t = __builtin_popcount(hi & lo & vm), g = __builtin_popcount(hi & ~lo & vm),
           c = __builtin_popcount(~hi & lo & vm);
#endif
  uint32_t a = (uint32_t)m - t - g - c;
  return a | c << 8 | g << 16 | t << 24;
}

// Scalar rank of all four symbols at k (inclusive), k in "with-$" coordinates; k == -1 -> zeros.
B2_HD inline void b2_occ4(const B2Index& ix, uint64_t k, uint64_t cnt[4]) {
  if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
  k -= (k >= ix.primary);
  const uint32_t* p = ix.bwt + ((k >> 7) << 4);
  const uint64_t* c = (const uint64_t*)p;
  int n = (int)(k & 127) + 1;
  uint32_t pc = 0;
  for (int wi = 0; wi < 8; ++wi) {
    int m = n - 16 * wi;
    if (m <= 0) break;
    pc += b2_cnt_word(p[8 + wi], m > 16 ? 16 : m);
  }
  cnt[0] = c[0] + (pc & 0xff); cnt[1] = c[1] + (pc >> 8 & 0xff);
  cnt[2] = c[2] + (pc >> 16 & 0xff); cnt[3] = c[3] + (pc >> 24);
}

// Scalar rank of one symbol (used by the SA walk).
B2_HD inline uint64_t b2_occ1(const B2Index& ix, uint64_t k, int c) {
  if (k == ix.seq_len) return ix.L2[c + 1] - ix.L2[c];
  if (k == (uint64_t)-1) return 0;
  k -= (k >= ix.primary);
  const uint32_t* p = ix.bwt + ((k >> 7) << 4);
  uint64_t n0 = ((const uint64_t*)p)[c];
  int n = (int)(k & 127) + 1;
  uint32_t pc = 0;
  for (int wi = 0; wi < 8; ++wi) {
    int m = n - 16 * wi;
    if (m <= 0) break;
    pc += b2_cnt_word(p[8 + wi], m > 16 ? 16 : m);
  }
  return n0 + (pc >> (8 * c) & 0xff);
}

B2_HD inline int b2_bwt_sym(const B2Index& ix, uint64_t x) {
  // symbol x of the $-removed BWT string
  uint32_t w = ix.bwt[((x >> 7) << 4) + 8 + ((x & 0x7f) >> 4)];
  return (w >> ((~x & 0xf) << 1)) & 3;
}

// One LF step (inverse Psi), and the sampled-SA walk.
B2_HD inline uint64_t b2_inv_psi(const B2Index& ix, uint64_t k) {
  if (k == ix.primary) return 0;
  uint64_t x = k - (k > ix.primary);
  int c = b2_bwt_sym(ix, x);
  return ix.L2[c] + b2_occ1(ix, k, c);
}

B2_HD inline uint64_t b2_sa(const B2Index& ix, uint64_t k, unsigned* n_steps = 0) {
  uint64_t steps = 0, mask = (uint64_t)ix.sa_intv - 1;
  while (k & mask) { ++steps; k = b2_inv_psi(ix, k); }
  if (n_steps) *n_steps = (unsigned)steps;
  return steps + ix.sa[k >> ix.sa_shift];
}

// ---------------------------------------------------------------------------------------------
// Rank pair provider: scalar (host checker builds / single-lane device code) or group-cooperative.
// 64-byte block touches counted the way SURVEY.md 8(d) defines the algorithmic seeding bytes: one per
// bwt_occ4 call with k != -1, one for the same-block path of bwt_2occ4 (bwa/bwt.c:177,200).
B2_HD inline unsigned b2_pair_blocks(const B2Index& ix, uint64_t k, uint64_t l) {
  const bool kv = (k != (uint64_t)-1), lv = (l != (uint64_t)-1);
  uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
  if (kv && lv && (kk >> 7) == (ll >> 7)) return 1;
  return (unsigned)kv + (unsigned)lv;
}

// packed counts (A | C<<8 | G<<16 | T<<24) of the top m (0..32) symbols of 32 symbols held MSB-first in a u64
B2_HD inline uint32_t b2_cnt_dword(uint64_t w, int m) {
  if (m <= 0) return 0;
  const uint64_t vm = 0x5555555555555555ull & (m >= 32 ? ~0ull : (~0ull << (64 - 2 * m)));
  const uint64_t hi = w >> 1, lo = w;
#if defined(__CUDA_ARCH__)
  uint32_t t = __popcll(hi & lo & vm), g = __popcll(hi & ~lo & vm), c = __popcll(~hi & lo & vm);
#else
  uint32_t t = __builtin_popcountll(hi & lo & vm), g = __builtin_popcountll(hi & ~lo & vm),
           c = __builtin_popcountll(~hi & lo & vm);
#endif
  uint32_t a = (uint32_t)(m > 32 ? 32 : m) - t - g - c;
  return a | c << 8 | g << 16 | t << 24;
}

// Child interval of `ik` for symbol c (one step of bwt_extend, bwa/bwt.c:262-275, keeping only ok[c]):
//   ok[c].x[f]  = L2[c] + 1 + rank_c(k)            k = ik.x[f]-1, l = k + ik.x[2], f = !is_back
//   ok[c].x[2]  = rank_c(l) - rank_c(k)
//   ok[c].x[b]  = ik.x[b] + [$ inside ik] + sum_{j>c} ok[j].x[2]
struct B2OccScalar {
  mutable unsigned long long n_blk = 0;
  B2_HD inline bool leader() const { return true; }
  B2_HD inline void child(const B2Index& ix, const B2Intv& ik, int c, int is_back, B2Intv& out) const {
    const int f = !is_back;
    uint64_t tk[4], tl[4];
    const uint64_t k = ik.x[f] - 1, l = ik.x[f] - 1 + ik.x[2];
    n_blk += b2_pair_blocks(ix, k, l);
    b2_occ4(ix, k, tk);
    b2_occ4(ix, l, tl);
    uint64_t xb = ik.x[is_back] + (ik.x[f] <= ix.primary && ik.x[f] + ik.x[2] - 1 >= ix.primary);
    for (int j = 3; j > c; --j) xb += tl[j] - tk[j];
    out.x[f] = ix.L2[c] + 1 + tk[c];
    out.x[2] = tl[c] - tk[c];
    out.x[is_back] = xb;
  }
};

#if defined(__CUDACC__)
// Group-cooperative form: lane g (0..3) of the group fetches, for each of the two blocks, the cumulative count of
// symbol g (8 bytes) and BWT words 2g, 2g+1 (8 bytes = 32 symbols), popcounts its 32 symbols for all four
// symbols, the packed partial counts are summed over the group with two butterfly steps, and lane g ends up
// with rank_g(k), rank_g(l).  A suffix sum over the lanes gives x[b]; the child of symbol c is broadcast from lane c.
struct B2OccCoop {
  unsigned mask;  // lanes of this group inside the warp
  int g;          // lane index inside the group (0..B2_SEED_LANES-1)
  mutable unsigned long long n_blk = 0;
  __device__ inline bool leader() const { return g == 0; }
  __device__ inline uint64_t bc64(uint64_t v, int src) const {
    return (uint64_t)__shfl_sync(mask, (unsigned)v, src, B2_SEED_LANES) |
           (uint64_t)__shfl_sync(mask, (unsigned)(v >> 32), src, B2_SEED_LANES) << 32;
  }
  __device__ inline void child(const B2Index& ix, const B2Intv& ik, int c, int is_back, B2Intv& out) const {
    const int f = !is_back;
    const uint64_t k = ik.x[f] - 1, l = ik.x[f] - 1 + ik.x[2];
    n_blk += b2_pair_blocks(ix, k, l);
    const bool kv = (k != (uint64_t)-1), lv = (l != (uint64_t)-1);
    const uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
    uint64_t ck = 0, cl = 0, wk = 0, wl = 0;
    if (kv) {
      const uint64_t* b = reinterpret_cast<const uint64_t*>(ix.bwt + ((kk >> 7) << 4));
      ck = __ldg(b + g);
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(b + 4 + g));
      wk = (uint64_t)w.x << 32 | w.y;
    }
    if (lv) {
      const uint64_t* b = reinterpret_cast<const uint64_t*>(ix.bwt + ((ll >> 7) << 4));
      cl = __ldg(b + g);
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(b + 4 + g));
      wl = (uint64_t)w.x << 32 | w.y;
    }
    uint32_t pk = kv ? b2_cnt_dword(wk, (int)(kk & 127) + 1 - 32 * g) : 0;
    uint32_t pl = lv ? b2_cnt_dword(wl, (int)(ll & 127) + 1 - 32 * g) : 0;
    pk += __shfl_xor_sync(mask, pk, 1, B2_SEED_LANES); pl += __shfl_xor_sync(mask, pl, 1, B2_SEED_LANES);
    pk += __shfl_xor_sync(mask, pk, 2, B2_SEED_LANES); pl += __shfl_xor_sync(mask, pl, 2, B2_SEED_LANES);
    const uint64_t tk = ck + (pk >> (8 * g) & 0xff), tl = cl + (pl >> (8 * g) & 0xff);  // ranks of symbol g
    const uint64_t x2 = tl - tk;
    // suffix sum of the child sizes over symbols > g
    uint64_t suf = x2;
    {
      uint64_t o = bc64(suf, (g + 1) & 3); if (g + 1 < 4) suf += o;
      o = bc64(suf, (g + 2) & 3); if (g + 2 < 4) suf += o;
    }
    const uint64_t xb = ik.x[is_back] + (ik.x[f] <= ix.primary && ik.x[f] + ik.x[2] - 1 >= ix.primary) + (suf - x2);
    out.x[f] = bc64(ix.L2[g] + 1 + tk, c);
    out.x[2] = bc64(x2, c);
    out.x[is_back] = bc64(xb, c);
  }
};
#endif

B2_HD inline void b2_set_intv(const B2Index& ix, int c, B2Intv& ik) {
  ik.x[0] = ix.L2[c] + 1;
  ik.x[2] = ix.L2[c + 1] - ix.L2[c];
  ik.x[1] = ix.L2[3 - c] + 1;
  ik.info = 0;
}

B2_HD inline void b2_reverse_intvs(B2Intv* a, int n) {
  for (int j = 0; j < n >> 1; ++j) { B2Intv t = a[n - 1 - j]; a[n - 1 - j] = a[j]; a[j] = t; }
}

// All SMEMs covering position x with interval size >= min_intv.  mem/ta/tb each hold len+1 entries.
// Returns the query position where the next search starts.
template <class Occ>
B2_HD inline int b2_smem1(const B2Index& ix, const Occ& occ, int len, const uint8_t* q, int x, int min_intv,
                          B2Intv* mem, int* n_mem, B2Intv* ta, B2Intv* tb) {
  *n_mem = 0;
  if (q[x] > 3) return x + 1;
  if (min_intv < 1) min_intv = 1;
  B2Intv ik, okc;
  okc.x[0] = okc.x[1] = okc.x[2] = okc.info = 0;
  B2Intv *prev = ta, *curr = tb;
  int n_prev, n_curr = 0, i;
  b2_set_intv(ix, q[x], ik);
  ik.info = (uint64_t)(x + 1);
  for (i = x + 1; i < len; ++i) {  // forward: longest match starting at x, recording size changes
    if (q[i] < 4) {
      int c = 3 - q[i];
      occ.child(ix, ik, c, 0, okc);
      if (okc.x[2] != ik.x[2]) {
        curr[n_curr++] = ik;
        if (okc.x[2] < (uint64_t)min_intv) break;
      }
      ik = okc;
      ik.info = (uint64_t)(i + 1);
    } else {
      curr[n_curr++] = ik;
      break;
    }
  }
  if (i == len) curr[n_curr++] = ik;
  b2_reverse_intvs(curr, n_curr);  // longest matches first
  int ret = (int)curr[0].info;
  { B2Intv* t = curr; curr = prev; prev = t; }
  n_prev = n_curr;
  int nm = 0;
  for (i = x - 1; i >= -1; --i) {  // backward: shrink the candidate set one base at a time
    int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
    n_curr = 0;
    for (int j = 0; j < n_prev; ++j) {
      const B2Intv& p = prev[j];
      if (c >= 0) occ.child(ix, p, c, 1, okc);
      if (c < 0 || okc.x[2] < (uint64_t)min_intv) {
        if (n_curr == 0) {
          if (nm == 0 || (uint64_t)(i + 1) < (mem[nm - 1].info >> 32)) {
            ik = p;
            ik.info |= (uint64_t)(i + 1) << 32;
            mem[nm++] = ik;
          }
        }
      } else if (n_curr == 0 || okc.x[2] != curr[n_curr - 1].x[2]) {
        okc.info = p.info;
        curr[n_curr++] = okc;
      }
    }
    if (n_curr == 0) break;
    { B2Intv* t = curr; curr = prev; prev = t; }
    n_prev = n_curr;
  }
  b2_reverse_intvs(mem, nm);
  *n_mem = nm;
  return ret;
}

template <class Occ>
B2_HD inline int b2_seed_strategy1(const B2Index& ix, const Occ& occ, int len, const uint8_t* q, int x, int min_len,
                                   int max_intv, B2Intv* m) {
  m->x[0] = m->x[1] = m->x[2] = m->info = 0;
  if (q[x] > 3) return x + 1;
  B2Intv ik, okc;
  okc.info = 0;
  b2_set_intv(ix, q[x], ik);
  for (int i = x + 1; i < len; ++i) {
    if (q[i] < 4) {
      int c = 3 - q[i];
      occ.child(ix, ik, c, 0, okc);
      if (okc.x[2] < (uint64_t)max_intv && i - x >= min_len) {
        *m = okc;
        m->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
        return i + 1;
      }
      ik = okc;
    } else return i + 1;
  }
  return len;
}

// Three seeding passes + sort by (start,end).  `out` has `cap` entries; returns the number of
// intervals found (may exceed cap: the caller treats that as a capacity overflow and retries).
// scratch: 3*(len+1) B2Intv.
template <class Occ>
B2_HD inline int b2_collect_intv(const B2Opt& opt, const B2Index& ix, const Occ& occ, int len, const uint8_t* seq,
                                 B2Intv* out, int cap, B2Intv* scratch) {
  B2Intv *mem1 = scratch, *ta = scratch + (len + 1), *tb = scratch + 2 * (len + 1);
  int n = 0, n1, x = 0;
  const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
  while (x < len) {  // pass 1: all SMEMs
    if (seq[x] < 4) {
      x = b2_smem1(ix, occ, len, seq, x, 1, mem1, &n1, ta, tb);
      for (int i = 0; i < n1; ++i) {
        int slen = (int)((uint32_t)mem1[i].info - (uint32_t)(mem1[i].info >> 32));
        if (slen >= opt.min_seed_len) { if (n < cap) out[n] = mem1[i]; ++n; }
      }
    } else ++x;
  }
  const int old_n = n < cap ? n : cap;
  for (int k = 0; k < old_n; ++k) {  // pass 2: re-seed inside long, rare SMEMs
    B2Intv p = out[k];
    int start = (int)(p.info >> 32), end = (int32_t)p.info;
    if (end - start < split_len || p.x[2] > (uint64_t)opt.split_width) continue;
    b2_smem1(ix, occ, len, seq, (start + end) >> 1, (int)(p.x[2] + 1), mem1, &n1, ta, tb);
    for (int i = 0; i < n1; ++i)
      if ((int)((uint32_t)mem1[i].info - (uint32_t)(mem1[i].info >> 32)) >= opt.min_seed_len) {
        if (n < cap) out[n] = mem1[i];
        ++n;
      }
  }
  if (opt.max_mem_intv > 0) {  // pass 3: LAST-like forward seeds
    x = 0;
    while (x < len) {
      if (seq[x] < 4) {
        B2Intv m;
        x = b2_seed_strategy1(ix, occ, len, seq, x, opt.min_seed_len, (int)opt.max_mem_intv, &m);
        if (m.x[2] > 0) { if (n < cap) out[n] = m; ++n; }
      } else ++x;
    }
  }
  // sort by info; equal keys denote the same substring hence identical intervals (any order is exact)
  int m = n < cap ? n : cap;
  for (int i = 1; i < m; ++i) {
    B2Intv t = out[i];
    int j = i - 1;
    while (j >= 0 && out[j].info > t.info) { out[j + 1] = out[j]; --j; }
    out[j + 1] = t;
  }
  return n;
}
