// Inter-task global alignment: one ksw_global2 (bwa/ksw.c:504-606) per LANE, the band held in registers by
// diagonal, ahead of the finalisation kernel.
//
// mem_reg2aln (bwa/bwamem.c:1057-1127) aligns a region between fixed end points; for short reads the band is a few
// cells wide (w = |length difference| + 3 is typical), so a row of the matrix is far too little work for a lane
// group and the per-row scan/shuffle overhead dominated the finalisation kernel (ncu, profiles/r1c).  Here every
// lane owns a whole alignment: diagonal d = j - i + w lives in register d, the diagonal predecessor H[i-1][j-1] is
// the same register, E[i][j] is register d+1 of the previous row, F runs left to right exactly as in the reference
// loop -- no communication at all, and the 32 lanes of a warp run 32 alignments in lock step (tasks are bucketed
// by band class so that the lanes of a warp execute the same unrolled row).  Direction codes are packed 5 per word
// and stored lane-interleaved, so the warp's stores coalesce.
//
// The results are SPECULATIVE: tasks are planned from the regions as they stand after the extension stage (first
// attempt of mem_reg2aln's band-doubling loop) and kept in a hash table keyed by (read, qb, qe, rb, re, band); the
// finalisation code looks a region up there and falls back to its own DP on a miss (rescued mates, patched
// regions, retries with a wider band, wide bands), so the output cannot depend on what was planned.
#pragma once
#include "b2_types.h"
#include "b2_ksw.cuh"
#include "b2_coopreg.cuh"

#define B2_GALN_MAXCIG 12
#define B2_GALN_CLASSES 4
#define B2_GALN_MAXREGS 2   // regions per read planned ahead (they are sorted by score)

struct B2GalnTask {
  int64_t seq_off;  // offset of the read in the batch's base array (identifies the read)
  int64_t rb, re;
  int32_t read, qb, qe;
  int32_t w_in;     // band handed to bwa_gen_cigar2 (part of the key)
  int32_t w;        // band of ksw_global2 after bwa_gen_cigar2's adjustment (bwa/bwa.c:141-150)
};

struct B2GalnRes {
  int32_t score;
  int32_t n_cigar;  // < 0: not available (the consumer computes the alignment itself)
  uint32_t cigar[B2_GALN_MAXCIG];
};

struct B2GalnCache {
  const B2GalnTask* tasks;
  const B2GalnRes* res;
  const int32_t* table;   // open addressing, -1 = empty
  uint32_t mask;          // table size - 1 (0: no cache)
  const uint8_t* seq_base;
};

B2_HD inline uint32_t b2_galn_hash(int64_t seq_off, int qb, int qe, int64_t rb, int64_t re, int w_in) {
  uint64_t h = (uint64_t)seq_off * 0x9E3779B97F4A7C15ull;
  h ^= (uint64_t)rb + 0x7F4A7C15ull + (h << 6) + (h >> 2);
  h ^= (uint64_t)re * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= ((uint64_t)(uint32_t)qb << 40) ^ ((uint64_t)(uint32_t)qe << 20) ^ (uint64_t)(uint32_t)w_in;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  return (uint32_t)h;
}

B2_HD inline bool b2_galn_same(const B2GalnTask& t, int64_t seq_off, int qb, int qe, int64_t rb, int64_t re, int w_in) {
  return t.seq_off == seq_off && t.qb == qb && t.qe == qe && t.rb == rb && t.re == re && t.w_in == w_in;
}

B2_HD inline const B2GalnRes* b2_galn_lookup(const B2GalnCache* gc, const uint8_t* query, int qb, int qe, int64_t rb,
                                             int64_t re, int w_in) {
  if (gc == 0 || gc->mask == 0) return 0;
  const int64_t seq_off = (int64_t)(query - gc->seq_base);
  uint32_t h = b2_galn_hash(seq_off, qb, qe, rb, re, w_in) & gc->mask;
  for (uint32_t probe = 0; probe <= gc->mask; ++probe, h = (h + 1) & gc->mask) {
    const int32_t t = gc->table[h];
    if (t < 0) return 0;
    if (b2_galn_same(gc->tasks[t], seq_off, qb, qe, rb, re, w_in)) return gc->res[t].n_cigar >= 0 ? &gc->res[t] : 0;
  }
  return 0;
}

// band class of a ksw_global2 band: 2w+1 diagonals in 9 / 17 / 33 / 65 registers; -1: too wide for this path
B2_HD inline int b2_galn_class(int w) { return w <= 4 ? 0 : (w <= 8 ? 1 : (w <= 16 ? 2 : (w <= 32 ? 3 : -1))); }
#define B2_GALN_MAX_ZWORDS 13  // direction words per row of the widest class

// One alignment, entirely lane-local.  query/target: base codes (already reversed for the reverse strand, as
// bwa_gen_cigar2 does); zw: direction words of this lane, word k of row i at zw[(i * NWZ + k) * zstride].
// Preconditions: the substitution matrix has bwa_fill_scmat's shape, 2w+1 <= ND, |qlen - tlen| <= w, qlen, tlen >= 1.
template <int ND>
B2_HD inline void b2_galn_one(int qlen, const uint8_t* query, int tlen, const uint8_t* target, const int8_t* mat, int o_del,
                              int e_del, int o_ins, int e_ins, int w, uint32_t* zw, int zstride, B2GalnRes* res,
                              unsigned long long* cells = 0) {
  constexpr int NWZ = (ND + 4) / 5;
  constexpr int NQ = (ND + 9) / 10;
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int sa = mat[0], sb = mat[1], sn = mat[4];
  const int nd = 2 * w + 1;
  int hp[ND], en[ND];
  uint32_t qw[NQ];  // 3-bit codes of the columns under the diagonals of the current row
B2_UNROLL
  for (int k = 0; k < NQ; ++k) qw[k] = 0;
B2_UNROLL
  for (int d = 0; d < ND; ++d) {
    const int j = d - w;
    hp[d] = (j >= 1 && j <= w) ? -(o_ins + e_ins * j) : B2_MINUS_INF;
    en[d] = B2_MINUS_INF;
    qw[d / 10] |= (uint32_t)((j >= 0 && j < qlen && d < nd) ? query[j] : 4) << (3 * (d % 10));
  }
  int last_h = 0;
  const int d_fin = qlen - tlen + w;
  for (int i = 0; i < tlen; ++i) {
    const int t = target[i];
    const int dmin = w > i ? w - i : 0;
    const int dend = qlen - i + w < nd ? qlen - i + w : nd;
    const int hb = i ? -(o_del + e_del * i) : 0;
    int f = B2_MINUS_INF;
    uint32_t zword = 0;
    if (cells) *cells += (unsigned long long)(dend - dmin);  // = end - beg of the reference's row (bwa/ksw.c:530-531)
B2_UNROLL
    for (int d = 0; d < ND; ++d) {
      if (d >= dmin && d < dend) {
        const int q = (int)(qw[d / 10] >> (3 * (d % 10))) & 7;
        const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
        const int m = (d == dmin && i <= w ? hb : hp[d]) + s;
        int e = d + 1 < ND ? en[d + 1] : B2_MINUS_INF;
        uint32_t dir = m >= e ? 0 : 1;
        int h = m >= e ? m : e;
        dir = h >= f ? dir : 2;
        h = h >= f ? h : f;
        int t2 = m - oe_del;
        e -= e_del;
        dir |= e > t2 ? 1 << 2 : 0;
        e = e > t2 ? e : t2;
        t2 = m - oe_ins;
        f -= e_ins;
        dir |= f > t2 ? 2 << 4 : 0;
        f = f > t2 ? f : t2;
        hp[d] = h; en[d] = e;
        if (d == d_fin) last_h = h;
        zword |= dir << (6 * (d % 5));
      }
      if (d % 5 == 4 || d == ND - 1) { zw[((long)i * NWZ + d / 5) * zstride] = zword; zword = 0; }
    }
    // next row: diagonal d sits one column further right
    const int jn = i + 1 - w + nd - 1;
    const uint32_t qin = (jn >= 0 && jn < qlen) ? query[jn] : 4;
B2_UNROLL
    for (int k = 0; k < NQ; ++k) qw[k] = (qw[k] >> 3) | ((k + 1 < NQ ? (qw[k + 1] & 7u) : 0u) << 27);
    {  // the entering column belongs to diagonal nd-1
      const int dl = nd - 1;
B2_UNROLL
      for (int k = 0; k < NQ; ++k)
        if (k == dl / 10) qw[k] = (qw[k] & ~(7u << (3 * (dl % 10)))) | (qin << (3 * (dl % 10)));
    }
  }
  res->score = last_h;
  // traceback (bwa/ksw.c:585-604); direction of cell (i, k): diagonal d = k - i + w
  uint32_t cg[B2_GALN_MAXCIG];
  int n = 0, which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
  bool ok = true;
  while (i >= 0 && k >= 0) {
    const int d = k - i + w;
    const uint32_t zv = zw[((long)i * NWZ + d / 5) * zstride] >> (6 * (d % 5));
    which = (int)(zv >> (which << 1)) & 3;
    const int op = which == 0 ? 0 : (which == 1 ? 2 : 1);
    if (n == 0 || op != (int)(cg[n - 1] & 0xf)) {
      if (n == B2_GALN_MAXCIG) { ok = false; break; }
      cg[n++] = 1u << 4 | (uint32_t)op;
    } else cg[n - 1] += 1u << 4;
    if (which == 0) { --i; --k; } else if (which == 1) --i; else --k;
  }
  if (ok && i >= 0) {
    if (n && (cg[n - 1] & 0xf) == 2) cg[n - 1] += (uint32_t)(i + 1) << 4;
    else if (n == B2_GALN_MAXCIG) ok = false;
    else cg[n++] = (uint32_t)(i + 1) << 4 | 2;
  }
  if (ok && k >= 0) {
    if (n && (cg[n - 1] & 0xf) == 1) cg[n - 1] += (uint32_t)(k + 1) << 4;
    else if (n == B2_GALN_MAXCIG) ok = false;
    else cg[n++] = (uint32_t)(k + 1) << 4 | 1;
  }
  if (!ok) { res->n_cigar = -1; return; }
  for (int a = 0; a < n; ++a) res->cigar[a] = cg[n - 1 - a];
  res->n_cigar = n;
}

// ---------------------------------------------------------------------------------------------
// Band arithmetic shared with the consumers (so that plan and lookup cannot drift apart)

// band of ksw_global2 chosen by bwa_gen_cigar2 (bwa/bwa.c:141-150)
B2_HD inline int b2_cigar_band(const B2Opt& opt, int l_query, int rlen, int w_) {
  const int8_t* mat = opt.mat;
  int w, max_gap, max_ins, max_del, min_w;
  max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - opt.o_ins) / opt.e_ins + 1.);
  max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - opt.o_del) / opt.e_del + 1.);
  max_gap = max_ins > max_del ? max_ins : max_del;
  max_gap = max_gap > 1 ? max_gap : 1;
  int dl = rlen - l_query; dl = dl < 0 ? -dl : dl;
  w = (max_gap + dl + 1) >> 1;
  w = w < w_ ? w : w_;
  min_w = dl + 3;
  w = w > min_w ? w : min_w;
  return w;
}
