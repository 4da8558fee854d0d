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
// On the device one lane serves one read (32 independent searches per warp).  The search is written as a
// state machine (B2SeedFsm below) so that every lane of a warp reaches the one expensive step -- a
// bidirectional extension: two 64-byte block fetches and their popcounts -- at the same program point:
// the nested loops of the reference would leave the 32 lanes in 32 different places.
#pragma once
#include "b2_types.h"

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
// Per-lane form used by k_seed (one read per lane, 32 independent searches per warp): the lane fetches the whole
// 64-byte block with four 16-byte loads that bypass L1 allocation (the blocks are touched once; L1 is kept for the
// read's own bases and the interval stack spill), and ranks all four symbols with 12 popcounts: the 128 symbols are
// handled as four word pairs whose high/low bit planes are interleaved into one 32-bit lane each.
struct B2Blk { uint4 c01, c23, w03, w47; };  // counts A,C | counts G,T | BWT words 0..3 | BWT words 4..7

// 32-byte load that bypasses L1 allocation (sm_100: LDG.256): half the memory-pipe work of two 16-byte loads when
// every lane addresses its own line
__device__ inline void b2_ldg_stream32(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

__device__ inline void b2_ld_blk(const uint32_t* p, B2Blk& b) {
  b2_ldg_stream32(p, b.c01, b.c23);
  b2_ldg_stream32(p + 8, b.w03, b.w47);
}

// counts of T, G, C among the top symbols of a word pair: ma2/mb2 = 2 * (symbols counted in word a / b), 0..32
__device__ inline void b2_pair_cnt(uint32_t wa, uint32_t wb, int n2, uint32_t& T, uint32_t& G, uint32_t& C) {
  const int ma2 = __vimin_s32_relu(n2, 32), mb2 = __vimin_s32_relu(n2 - 32, 32);
  const uint32_t sa = ~__funnelshift_rc(0xFFFFFFFFu, 0u, ma2), sb = ~__funnelshift_rc(0xFFFFFFFFu, 0u, mb2);
  const uint32_t K = (sa & 0x55555555u) | (sb & 0xAAAAAAAAu);
  const uint32_t X = ((wa >> 1) & 0x55555555u) | (wb & 0xAAAAAAAAu);   // high bits: word a on even, word b on odd positions
  const uint32_t Y = (wa & 0x55555555u) | ((wb << 1) & 0xAAAAAAAAu);   // low bits, same positions
  T += __popc(X & Y & K); G += __popc(X & ~Y & K); C += __popc(~X & Y & K);
}

// ranks of A,C,G,T over symbols [0, n) of the block, n in 1..128, added to the block's cumulative counts
__device__ inline void b2_rank4(const B2Blk& b, int n, uint64_t r[4]) {
  uint32_t T = 0, G = 0, C = 0;
  const int n2 = 2 * n;
  b2_pair_cnt(b.w03.x, b.w03.y, n2, T, G, C);
  b2_pair_cnt(b.w03.z, b.w03.w, n2 - 64, T, G, C);
  b2_pair_cnt(b.w47.x, b.w47.y, n2 - 128, T, G, C);
  b2_pair_cnt(b.w47.z, b.w47.w, n2 - 192, T, G, C);
  r[0] = ((uint64_t)b.c01.y << 32 | b.c01.x) + ((uint32_t)n - T - G - C);
  r[1] = ((uint64_t)b.c01.w << 32 | b.c01.z) + C;
  r[2] = ((uint64_t)b.c23.y << 32 | b.c23.x) + G;
  r[3] = ((uint64_t)b.c23.w << 32 | b.c23.z) + T;
}

// One LF step (bwt_invPsi, bwa/bwt.c:53-59) for a lane of k_sa.  The 64-byte block that holds row k gives both the BWT
// symbol at k and the rank of that symbol up to and including k, so the step is ONE fetch (two 32-byte loads issued
// together) where symbol-then-count would be two dependent round trips to L2/HBM; and the rank is taken for that one
// symbol only, branch-free, so that the 32 lanes of a warp -- each at its own row -- run the same instructions:
// the block's words are XORed with the symbol replicated into every 2-bit field, after which the wanted symbol reads 11
// and one AND of the two bit planes per word pair marks it.
__device__ inline uint32_t b2_cnt1_pair(uint32_t wa, uint32_t wb, int n2) {
  const int ma2 = __vimin_s32_relu(n2, 32), mb2 = __vimin_s32_relu(n2 - 32, 32);
  const uint32_t sa = ~__funnelshift_rc(0xFFFFFFFFu, 0u, ma2), sb = ~__funnelshift_rc(0xFFFFFFFFu, 0u, mb2);
  const uint32_t ta = wa & (wa >> 1) & sa;   // marker on the low bit of a field of word a ...
  const uint32_t tb = wb & (wb << 1) & sb;   // ... and on the high bit of a field of word b
  return __popc((ta & 0x55555555u) | (tb & 0xAAAAAAAAu));
}

__device__ inline uint64_t b2_inv_psi_lane(const B2Index& ix, uint64_t k) {
  if (k == ix.primary) return 0;
  const uint64_t x = k - (k > ix.primary);   // row in the $-less BWT string
  B2Blk b;
  b2_ld_blk(ix.bwt + ((x >> 7) << 4), b);
  const int off = (int)(x & 127), wi = off >> 4;
  const uint32_t lo4 = (wi & 2) ? ((wi & 1) ? b.w03.w : b.w03.z) : ((wi & 1) ? b.w03.y : b.w03.x);
  const uint32_t hi4 = (wi & 2) ? ((wi & 1) ? b.w47.w : b.w47.z) : ((wi & 1) ? b.w47.y : b.w47.x);
  const uint32_t w = (wi & 4) ? hi4 : lo4;
  const int c = (int)(w >> ((~off & 15) << 1)) & 3;
  const uint32_t flip = ~((uint32_t)c * 0x55555555u);
  const int n2 = 2 * (off + 1);
  const uint32_t cnt = b2_cnt1_pair(b.w03.x ^ flip, b.w03.y ^ flip, n2) + b2_cnt1_pair(b.w03.z ^ flip, b.w03.w ^ flip, n2 - 64) +
                       b2_cnt1_pair(b.w47.x ^ flip, b.w47.y ^ flip, n2 - 128) + b2_cnt1_pair(b.w47.z ^ flip, b.w47.w ^ flip, n2 - 192);
  const uint32_t n_lo = (c & 2) ? ((c & 1) ? b.c23.z : b.c23.x) : ((c & 1) ? b.c01.z : b.c01.x);
  const uint32_t n_hi = (c & 2) ? ((c & 1) ? b.c23.w : b.c23.y) : ((c & 1) ? b.c01.w : b.c01.y);
  return ix.L2[c] + ((uint64_t)n_hi << 32 | n_lo) + cnt;
}

struct B2OccLane {
  unsigned long long n_blk = 0;
  __device__ inline void child(const B2Index& ix, const B2Intv& ik, int c, int back, B2Intv& out) {
    const uint64_t xf = back ? ik.x[0] : ik.x[1], xo = back ? ik.x[1] : ik.x[0];
    const uint64_t k = xf - 1, l = xf - 1 + ik.x[2];
    const bool kv = (k != (uint64_t)-1), lv = (l != (uint64_t)-1);
    const uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
    const bool same = kv && lv && (kk >> 7) == (ll >> 7);
    n_blk += same ? 1u : (unsigned)kv + (unsigned)lv;
    B2Blk bk, bl;
    bk.c01 = bk.c23 = bk.w03 = bk.w47 = make_uint4(0, 0, 0, 0);
    bl = bk;
    if (kv) b2_ld_blk(ix.bwt + ((kk >> 7) << 4), bk);
    if (lv && !same) b2_ld_blk(ix.bwt + ((ll >> 7) << 4), bl);
    if (same) bl = bk;
    uint64_t tk[4] = {0, 0, 0, 0}, tl[4] = {0, 0, 0, 0};
    if (kv) b2_rank4(bk, (int)(kk & 127) + 1, tk);
    if (lv) b2_rank4(bl, (int)(ll & 127) + 1, tl);
    const uint64_t d0 = tl[0] - tk[0], d1 = tl[1] - tk[1], d2 = tl[2] - tk[2], d3 = tl[3] - tk[3];
    const uint64_t dc = c == 0 ? d0 : c == 1 ? d1 : c == 2 ? d2 : d3;
    const uint64_t tc = c == 0 ? tk[0] : c == 1 ? tk[1] : c == 2 ? tk[2] : tk[3];
    const uint64_t suf = c == 0 ? d1 + d2 + d3 : c == 1 ? d2 + d3 : c == 2 ? d3 : 0;
    const uint64_t of = ix.L2[c] + 1 + tc;
    const uint64_t ob = xo + (xf <= ix.primary && xf + ik.x[2] - 1 >= ix.primary) + suf;
    out.x[0] = back ? of : ob;
    out.x[1] = back ? ob : of;
    out.x[2] = dc;
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

B2_HD inline void b2_sort_intv(B2Intv* a, int m);

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
  b2_sort_intv(out, n < cap ? n : cap);
  return n;
}

// ---------------------------------------------------------------------------------------------
// mem_collect_intv as a resumable state machine (one per lane).  consume() takes the child interval of the
// pending request, does the lane-local bookkeeping and leaves the next request (interval `p`, symbol `rc`,
// direction `rback`) or reports that the read is finished; the caller performs the extension for all lanes
// of the warp together.  What differs from the nested-loop form above, without changing the result:
//   * the work arrays of bwt_smem1a (`prev`/`curr`, bwa/bwt.c:296-343) are one interval stack: the forward
//     phase pushes entries 0..top-1, the backward sweeps address entry k as stack[top-1-k] (longest match
//     first, i.e. the reversed order) and compact in place (a sweep writes entry n_curr <= j after reading j);
//   * SMEMs are filtered by min_seed_len and appended to `out` as found; `out` is left unsorted -- k_seed_fin
//     sorts it by `info` (equal keys are identical intervals, so the sorted array does not depend on the
//     order of insertion) -- and the re-seeding pass takes its candidates (bwamem.c:133-142) from a short
//     list recorded when they are emitted rather than re-reading `out`;
//   * everything touched per step (stack, read bases, candidate list) is in shared memory: a warp runs the
//     32 lanes' divergent bookkeeping serially, so a global-memory access there would stall all of them.
enum { B2S_IDLE = 0, B2S_FWD, B2S_BWD, B2S_S3 };
enum { B2T_SCAN1 = 0, B2T_START, B2T_BEGIN_BWD, B2T_SWEEP_END, B2T_SWEEP_START, B2T_END_SMEM, B2T_SCAN2, B2T_SCAN3, B2T_DONE };

struct B2SeedMemHost {  // plain arrays (host checker build)
  B2Intv* stk; const uint8_t* qg; uint32_t* rsv;
  B2_HD inline void load_query(const uint8_t* q, int) { qg = q; }
  B2_HD inline int q(int i) const { return qg[i]; }
  B2_HD inline void get(int i, B2Intv& v) const { v = stk[i]; v.x[1] = 0; }  // (x[1] is not tracked through the backward phase)
  B2_HD inline void put(int i, const B2Intv& v) { stk[i] = v; }
  B2_HD inline void rs_put(int k, uint32_t x, uint32_t mi) { rsv[2 * k] = x; rsv[2 * k + 1] = mi; }
  B2_HD inline void rs_get(int k, uint32_t& x, uint32_t& mi) const { x = rsv[2 * k]; mi = rsv[2 * k + 1]; }
};

#if defined(__CUDACC__)
// Device: per lane E stack entries of three words {x0.lo, x2.lo, x0.hi | x2.hi << 4 | info << 8}: four high bits for each
// interval bound and 24 bits for the query end (the engine refuses indexes of 2^36 symbols or more -- a 34 Gbp
// reference -- and reads of 2^24 bases or more), QW words of
// 4-bit base codes and RS candidate pairs, all word-transposed over the block's lanes (bank-conflict free).
// Deeper stack entries, longer reads and further candidates spill to a per-lane global array.
// x[1] -- the interval of the reverse complement -- is not kept on the stack: backward extension only passes it
// through, and nothing after seeding reads it (mem_chain uses x[0], x[2] and info; bwamem.c:273-301), so SMEMs
// leave this kernel with x[1] = 0.  Forward walks (where x[1] drives the extension) keep it in registers.
// The spill paths are out-of-line calls on purpose: inlined, the compiler predicates them, and a predicated-off
// global load/store still occupies the memory pipe and a long-scoreboard slot in every trip of the hot loop.
__device__ __noinline__ void b2_spill_get3(const uint32_t* a, uint32_t* w) { w[0] = a[0]; w[1] = a[1]; w[2] = a[2]; }
__device__ __noinline__ void b2_spill_put3(uint32_t* a, uint32_t w0, uint32_t w1, uint32_t w2) { a[0] = w0; a[1] = w1; a[2] = w2; }
__device__ __noinline__ void b2_spill_put2(uint32_t* a, uint32_t w0, uint32_t w1) { a[0] = w0; a[1] = w1; }
__device__ __noinline__ int b2_spill_q(const uint8_t* q, int i) { return q[i]; }

struct B2SeedMemDev {
  uint32_t *s, *qs, *rs;   // shared: stack, bases, candidates (already offset by the lane)
  uint32_t *g, *grs;       // global spill: stack entries >= E, candidates >= RS
  const uint8_t* qg;
  int stride, E, QW, RS;
  __device__ inline void load_query(const uint8_t* q, int len) {
    qg = q;
    const int nw = (len + 7) >> 3 < QW ? (len + 7) >> 3 : QW;
    for (int w = 0; w < nw; ++w) {
      uint32_t v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int i = 8 * w + k; v |= (uint32_t)(i < len ? q[i] : 4) << (4 * k); }
      qs[w * stride] = v;
    }
  }
  __device__ inline int q(int i) const {
    if (i < 8 * QW) return (int)(qs[(i >> 3) * stride] >> ((i & 7) * 4) & 15);
    return b2_spill_q(qg, i);
  }
  __device__ inline void get(int i, B2Intv& v) const {
    uint32_t w[3];
    if (i < E) { const uint32_t* a = s + (size_t)i * 3 * stride; w[0] = a[0]; w[1] = a[stride]; w[2] = a[2 * stride]; }
    else b2_spill_get3(g + (size_t)(i - E) * 3, w);
    v.x[0] = (uint64_t)(w[2] & 0xf) << 32 | w[0];
    v.x[1] = 0;
    v.x[2] = (uint64_t)(w[2] >> 4 & 0xf) << 32 | w[1];
    v.info = w[2] >> 8;
  }
  __device__ inline void put(int i, const B2Intv& v) {
    const uint32_t w2 = (uint32_t)(v.x[0] >> 32) | (uint32_t)(v.x[2] >> 32) << 4 | (uint32_t)v.info << 8;
    if (i < E) { uint32_t* a = s + (size_t)i * 3 * stride; a[0] = (uint32_t)v.x[0]; a[stride] = (uint32_t)v.x[2]; a[2 * stride] = w2; }
    else b2_spill_put3(g + (size_t)(i - E) * 3, (uint32_t)v.x[0], (uint32_t)v.x[2], w2);
  }
  __device__ inline void rs_put(int k, uint32_t x, uint32_t mi) {
    if (k < RS) { rs[2 * k * stride] = x; rs[(2 * k + 1) * stride] = mi; } else b2_spill_put2(grs + 2 * (k - RS), x, mi);
  }
  __device__ inline void rs_get(int k, uint32_t& x, uint32_t& mi) const {
    if (k < RS) { x = rs[2 * k * stride]; mi = rs[(2 * k + 1) * stride]; }
    else { uint32_t w[3]; b2_spill_get3(grs + 2 * (k - RS), w); x = w[0]; mi = w[1]; }
  }
};
#endif

template <class Mem, int MODE>
struct B2SeedFsm {
  Mem mem;
  int len; B2Intv* out; int cap;
  int phase, pass;
  int x, i, j, n_prev, n_curr, top, n, nm_last, ret, k2, n_rs, rc, rback;
  uint64_t min_intv, last_x2;
  B2Intv ik, p;

  B2_HD inline void emit(const B2Opt& opt, const B2Intv& v, int start) {
    const int qend = (int)v.info;
    nm_last = start;
    if (qend - start >= opt.min_seed_len) {
      if (n < cap) { B2Intv t = v; t.info = (uint64_t)start << 32 | (uint32_t)qend; out[n] = t; }
      ++n;
      if (pass == 1 && qend - start >= (int)(opt.min_seed_len * opt.split_factor + .499) && v.x[2] <= (uint64_t)opt.split_width)
        mem.rs_put(n_rs++, (uint32_t)((start + qend) >> 1), (uint32_t)(v.x[2] + 1));
    }
  }

  // First request of a task; returns 0 if it yields nothing to extend.  A read is two independent tasks, run by
  // different lanes: mode 0 = SMEM passes 1+2 (entries out[0], out[1], ...), mode 1 = pass 3, the forward-only
  // seeds of bwt_seed_strategy1 (entries out[cap-1], out[cap-2], ...); k_seed_fin joins and sorts them.
  B2_HD inline int begin(const B2Opt& opt, const B2Index& ix, const uint8_t* q, int len_, B2Intv* out_, int cap_) {
    len = len_; out = out_; cap = cap_;
    mem.load_query(q, len);
    n = 0; n_rs = 0; x = 0; pass = MODE ? 3 : 1;
    if (MODE) return opt.max_mem_intv > 0 ? advance(opt, ix, B2T_SCAN3) : 0;
    return advance(opt, ix, B2T_SCAN1);
  }

  // lane-local transitions until the next request (returns 1) or the end of the read (returns 0)
  B2_HD inline int advance(const B2Opt& opt, const B2Index& ix, int t) {
    for (;;) {
      if (MODE == 0 && (t == B2T_START || t == B2T_SCAN1)) {
        if (t == B2T_SCAN1) {
          while (x < len && mem.q(x) > 3) ++x;
          if (x >= len) { pass = 2; k2 = 0; t = B2T_SCAN2; continue; }
          min_intv = 1;
        }
        b2_set_intv(ix, mem.q(x), ik);  // start of bwt_smem1a at x
        ik.info = (uint64_t)(x + 1);
        i = x + 1; n_curr = 0; phase = B2S_FWD; rback = 0;
        if (i < len) { const int b = mem.q(i); if (b < 4) { p = ik; rc = 3 - b; return 1; } }
        mem.put(n_curr++, ik); ret = (int)ik.info;
        t = B2T_BEGIN_BWD;
      } else if (MODE == 0 && (t == B2T_BEGIN_BWD || t == B2T_SWEEP_END || t == B2T_SWEEP_START)) {
        if (t == B2T_BEGIN_BWD) { top = n_curr; n_prev = n_curr; i = x - 1; nm_last = 0x7fffffff; phase = B2S_BWD; rback = 1; }
        else if (t == B2T_SWEEP_END) {
          if (n_curr == 0) { t = B2T_END_SMEM; continue; }
          n_prev = n_curr; --i;
        }
        j = 0; n_curr = 0;
        int bc = -1;
        if (i >= 0) { bc = mem.q(i); if (bc > 3) bc = -1; }
        mem.get(top - 1, p);
        if (bc < 0) {  // nothing extends: only the longest candidate can be reported
          if (i + 1 < nm_last) emit(opt, p, i + 1);
          t = B2T_END_SMEM;
          continue;
        }
        rc = bc;
        return 1;
      } else if (MODE == 0 && t == B2T_END_SMEM) {
        if (pass == 1) { x = ret; t = B2T_SCAN1; } else t = B2T_SCAN2;
      } else if (MODE == 0 && t == B2T_SCAN2) {
        if (k2 < n_rs) {
          uint32_t rx, rmi;
          mem.rs_get(k2++, rx, rmi);
          x = (int)rx; min_intv = rmi; t = B2T_START;
        } else t = B2T_DONE;
      } else if (MODE == 1 && t == B2T_SCAN3) {
        while (x < len && mem.q(x) > 3) ++x;
        if (x >= len) { t = B2T_DONE; continue; }
        b2_set_intv(ix, mem.q(x), ik);
        i = x + 1; phase = B2S_S3; rback = 0;
        if (i >= len) { t = B2T_DONE; continue; }
        const int b = mem.q(i);
        if (b < 4) { p = ik; rc = 3 - b; return 1; }
        x = i + 1;
      } else {
        phase = B2S_IDLE;
        return 0;
      }
    }
  }

  B2_HD inline int consume(const B2Opt& opt, const B2Index& ix, B2Intv okc) {
    int t;
    if (MODE == 0 && phase == B2S_BWD) {
      if (okc.x[2] < min_intv) {
        if (n_curr == 0 && i + 1 < nm_last) emit(opt, p, i + 1);
      } else if (n_curr == 0 || okc.x[2] != last_x2) {
        okc.info = p.info;
        mem.put(top - 1 - n_curr, okc);
        ++n_curr;
        last_x2 = okc.x[2];
      }
      ++j;
      // The end of a sweep (j == n_prev) is the COMMON case -- a sweep has one to three candidates -- so the
      // transition to the next sweep is taken here, in line with the lanes that stay inside their sweep, instead of
      // through advance() (whose state loop ran at ~5 of 32 lanes and made up 40 % of the kernel's instructions).
      if (j >= n_prev) {
        if (n_curr == 0) return advance(opt, ix, B2T_END_SMEM);
        n_prev = n_curr; --i;
        j = 0; n_curr = 0;
        int bc = -1;
        if (i >= 0) { bc = mem.q(i); if (bc > 3) bc = -1; }
        if (bc < 0) {  // nothing extends: only the longest candidate can be reported
          mem.get(top - 1, p);
          if (i + 1 < nm_last) emit(opt, p, i + 1);
          return advance(opt, ix, B2T_END_SMEM);
        }
        rc = bc;
      }
      mem.get(top - 1 - j, p);
      return 1;
    } else {  // one more base of a forward walk: bwt_smem1a's first loop (B2S_FWD) or bwt_seed_strategy1 (B2S_S3)
      const bool fwd = MODE == 0;
      bool stop;
      if (fwd) {
        const bool changed = okc.x[2] != ik.x[2];
        if (changed) { mem.put(n_curr++, ik); ret = (int)ik.info; }
        stop = changed && okc.x[2] < min_intv;
        t = B2T_BEGIN_BWD;
      } else {
        stop = okc.x[2] < opt.max_mem_intv && i - x >= opt.min_seed_len;
        if (stop && okc.x[2] > 0) {
          if (n < cap) { okc.info = (uint64_t)x << 32 | (uint64_t)(i + 1); out[cap - 1 - n] = okc; }
          ++n;
        }
        if (stop) x = i + 1;  // where the next walk starts
        t = B2T_SCAN3;
      }
      if (!stop) {
        ik = okc; ik.info = (uint64_t)(i + 1);
        ++i;
        if (i < len) {
          const int b = mem.q(i);
          if (b < 4) { p = ik; rc = 3 - b; return 1; }
        }
        if (fwd) { mem.put(n_curr++, ik); ret = (int)ik.info; }  // ambiguous base or end of the read
        else if (i >= len) t = B2T_DONE;
        else x = i + 1;
      }
    }
    return advance(opt, ix, t);
  }
};

// sort by info (insertion sort: a handful of entries); equal keys denote the same substring hence identical intervals
B2_HD inline void b2_sort_intv(B2Intv* a, int m) {
  for (int i = 1; i < m; ++i) {
    B2Intv t = a[i];
    int j = i - 1;
    while (j >= 0 && a[j].info > t.info) { a[j + 1] = a[j]; --j; }
    a[j + 1] = t;
  }
}
