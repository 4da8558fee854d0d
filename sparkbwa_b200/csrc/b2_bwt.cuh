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

struct B2OccScalar {
  mutable unsigned long long n_blk = 0;
  B2_HD inline void pair(const B2Index& ix, uint64_t k, uint64_t l, uint64_t tk[4], uint64_t tl[4]) const {
    n_blk += b2_pair_blocks(ix, k, l);
    b2_occ4(ix, k, tk);
    b2_occ4(ix, l, tl);
  }
  B2_HD inline bool leader() const { return true; }
};

#if defined(__CUDACC__)
struct B2OccCoop {
  unsigned mask;  // lanes of this group inside the warp
  int g;          // lane index inside the group (0..B2_SEED_LANES-1)
  mutable unsigned long long n_blk = 0;
  __device__ inline bool leader() const { return g == 0; }
  __device__ inline void pair(const B2Index& ix, uint64_t k, uint64_t l, uint64_t tk[4], uint64_t tl[4]) const {
    n_blk += b2_pair_blocks(ix, k, l);
    const bool kv = (k != (uint64_t)-1), lv = (l != (uint64_t)-1);
    uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
    uint4 vk = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
    if (kv) vk = __ldg(reinterpret_cast<const uint4*>(ix.bwt + ((kk >> 7) << 4)) + g);
    if (lv) vl = __ldg(reinterpret_cast<const uint4*>(ix.bwt + ((ll >> 7) << 4)) + g);
    uint32_t pk = 0, pl = 0;
    if (g >= 2) {
      int base = (g - 2) * 64;
      int nk = (int)(kk & 127) + 1 - base, nl = (int)(ll & 127) + 1 - base;
      const uint32_t wk[4] = {vk.x, vk.y, vk.z, vk.w}, wl[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        int mk = nk - 16 * wi, ml = nl - 16 * wi;
        pk += b2_cnt_word(wk[wi], mk > 16 ? 16 : mk);
        pl += b2_cnt_word(wl[wi], ml > 16 ? 16 : ml);
      }
    }
    pk = __shfl_sync(mask, pk, 2, B2_SEED_LANES) + __shfl_sync(mask, pk, 3, B2_SEED_LANES);
    pl = __shfl_sync(mask, pl, 2, B2_SEED_LANES) + __shfl_sync(mask, pl, 3, B2_SEED_LANES);
#define B2_BC64(v_lo, v_hi, src) \
  ((uint64_t)__shfl_sync(mask, v_lo, src, B2_SEED_LANES) | (uint64_t)__shfl_sync(mask, v_hi, src, B2_SEED_LANES) << 32)
    tk[0] = B2_BC64(vk.x, vk.y, 0) + (pk & 0xff);
    tk[1] = B2_BC64(vk.z, vk.w, 0) + (pk >> 8 & 0xff);
    tk[2] = B2_BC64(vk.x, vk.y, 1) + (pk >> 16 & 0xff);
    tk[3] = B2_BC64(vk.z, vk.w, 1) + (pk >> 24);
    tl[0] = B2_BC64(vl.x, vl.y, 0) + (pl & 0xff);
    tl[1] = B2_BC64(vl.z, vl.w, 0) + (pl >> 8 & 0xff);
    tl[2] = B2_BC64(vl.x, vl.y, 1) + (pl >> 16 & 0xff);
    tl[3] = B2_BC64(vl.z, vl.w, 1) + (pl >> 24);
#undef B2_BC64
    if (!kv) tk[0] = tk[1] = tk[2] = tk[3] = 0;
    if (!lv) tl[0] = tl[1] = tl[2] = tl[3] = 0;
  }
};
#endif

// One bidirectional extension step: children of `ik` for the four symbols.
template <class Occ>
B2_HD inline void b2_extend(const B2Index& ix, const Occ& occ, const B2Intv& ik, B2Intv ok[4], int is_back) {
  uint64_t tk[4], tl[4];
  const int f = !is_back;
  occ.pair(ix, ik.x[f] - 1, ik.x[f] - 1 + ik.x[2], tk, tl);
  for (int i = 0; i < 4; ++i) {
    ok[i].x[f] = ix.L2[i] + 1 + tk[i];
    ok[i].x[2] = tl[i] - tk[i];
  }
  // the '$' sentinel sorts before A in the complementary direction
  ok[3].x[is_back] = ik.x[is_back] + (ik.x[f] <= ix.primary && ik.x[f] + ik.x[2] - 1 >= ix.primary);
  ok[2].x[is_back] = ok[3].x[is_back] + ok[3].x[2];
  ok[1].x[is_back] = ok[2].x[is_back] + ok[2].x[2];
  ok[0].x[is_back] = ok[1].x[is_back] + ok[1].x[2];
}

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
  B2Intv ik, ok[4];
  B2Intv *prev = ta, *curr = tb;
  int n_prev, n_curr = 0, i;
  b2_set_intv(ix, q[x], ik);
  ik.info = (uint64_t)(x + 1);
  for (i = x + 1; i < len; ++i) {  // forward: longest match starting at x, recording size changes
    if (q[i] < 4) {
      int c = 3 - q[i];
      b2_extend(ix, occ, ik, ok, 0);
      if (ok[c].x[2] != ik.x[2]) {
        curr[n_curr++] = ik;
        if (ok[c].x[2] < (uint64_t)min_intv) break;
      }
      ik = ok[c];
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
      if (c >= 0) b2_extend(ix, occ, p, ok, 1);
      if (c < 0 || ok[c].x[2] < (uint64_t)min_intv) {
        if (n_curr == 0) {
          if (nm == 0 || (uint64_t)(i + 1) < (mem[nm - 1].info >> 32)) {
            ik = p;
            ik.info |= (uint64_t)(i + 1) << 32;
            mem[nm++] = ik;
          }
        }
      } else if (n_curr == 0 || ok[c].x[2] != curr[n_curr - 1].x[2]) {
        ok[c].info = p.info;
        curr[n_curr++] = ok[c];
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
  B2Intv ik, ok[4];
  b2_set_intv(ix, q[x], ik);
  for (int i = x + 1; i < len; ++i) {
    if (q[i] < 4) {
      int c = 3 - q[i];
      b2_extend(ix, occ, ik, ok, 0);
      if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
        *m = ok[c];
        m->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
        return i + 1;
      }
      ik = ok[c];
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
