// Inter-task seed extension: one ksw_extend2 (bwa/ksw.c:380-479) per LANE, ahead of the chain-extension kernel.
//
// The group-cooperative ksw_extend2 spends ~250 warp instructions per matrix row on scans, reductions and band
// bookkeeping for ~45 cells (ncu, profiles/r1c); a lane running the reference's plain row loop needs ~28 per cell
// and no communication.  Here the 32 lanes of a warp run 32 extensions row-synchronously.  The eh[] array of the
// reference ({H of the previous row, E}; both 0..65535 for short reads) is packed into one 32-bit word per column
// and kept in shared memory lane-interleaved (word j of lane l at j*32 + l: bank = l, conflict-free whatever column
// each lane is at); the two sequences are staged the same way as 4-bit codes.
//
// Like b2_galn.cuh the results are SPECULATIVE: per chain, the first band try of the left and of the right extension
// of the chain's best seed (the one mem_chain2aln always extends) are computed ahead and stored per seed;
// b2_extend_seed takes them when (band, start score) match and otherwise runs its own DP.
#pragma once
#include "b2_types.h"
#include "b2_ksw.cuh"
#include "b2_coopreg.cuh"

struct B2XextTask {
  int64_t tpos;      // bidirectional reference position of target[0]
  int64_t qsrc;      // offset in the batch's base array of query[0]
  int32_t slot;      // result slot: 2 * (global seed index) + side
  int32_t qlen, tlen;
  int32_t qdir, tdir;  // +1 / -1: direction in which query / target run through the read / the reference
  int32_t h0, w, end_bonus;
};

struct B2XextRes {
  int32_t w, h0;     // key: band and start score this result was computed for (w < 0: no result)
  int32_t score, qle, tle, gtle, gscore, max_off;
};

// element k of a lane-strided array.  Cell = uint32_t: {h, e} as 16 + 16 bits; Cell = uint16_t: 8 + 8 bits, for batches
// whose scores stay below 256 (read length x match score <= 250: every 2x150 bp and 250 bp workload at -A 1) -- half
// the shared memory per extension, i.e. nearly twice the warps per SM for a kernel that is latency-bound.
template <class Cell>
struct B2XMemT {
  Cell* eh;      // qlen+1 cells: h | e << (4 * sizeof(Cell))
  uint32_t* q;   // 4-bit codes, 8 per word
  uint32_t* t;
  int stride;
};
typedef B2XMemT<uint32_t> B2XMem;

// base code at bidirectional position p (forward strand [0, l_pac), reverse complement [l_pac, 2 l_pac))
B2_HD inline int b2_base_at(const uint8_t* pac, int64_t l_pac, int64_t p) {
  if (p < l_pac) return pac[p >> 2] >> ((~p & 3) << 1) & 3;
  const int64_t k = (l_pac << 1) - 1 - p;
  return 3 - (pac[k >> 2] >> ((~k & 3) << 1) & 3);
}

// Same results as ksw_extend2 (fuzzed against its scalar form in tests/coop_fuzz.cpp) for matrices of bwa_fill_scmat's shape and h0 + qlen*max(mat) < 65536.
template <class Cell>
B2_HD inline int b2_xext_one(const B2XMemT<Cell>& M, int qlen, int tlen, const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins,
                             int w, int end_bonus, int zdrop, int h0, int* _qle, int* _tle, int* _gtle, int* _gscore,
                             int* _max_off, unsigned long long* cells) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int sa = mat[0], sb = mat[1], sn = mat[4];
  const int st = M.stride;
  constexpr int ESH = 4 * (int)sizeof(Cell);            // bit position of e inside a cell
  constexpr uint32_t HMASK = (1u << ESH) - 1;
  int i, j, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
  {
    const int a1 = h0 > oe_ins ? h0 - oe_ins : 0;
    M.eh[0] = (Cell)h0;
    for (j = 1; j <= qlen; ++j) { int h = a1 - (j - 1) * e_ins; M.eh[(size_t)j * st] = (Cell)(h > 0 ? h : 0); }
  }
  max = 0;
  for (i = 0; i < 25; ++i) max = max > mat[i] ? max : mat[i];
  max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
  max_ins = max_ins > 1 ? max_ins : 1;
  w = w < max_ins ? w : max_ins;
  max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
  max_del = max_del > 1 ? max_del : 1;
  w = w < max_del ? w : max_del;
  max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
  beg = 0; end = qlen;
  unsigned long long ncell = 0;
  for (i = 0; i < tlen; ++i) {
    int f = 0, h1, m = 0, mj = -1;
    const int t = (int)(M.t[(size_t)(i >> 3) * st] >> ((i & 7) << 2)) & 15;
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    ncell += (unsigned long long)(end > beg ? end - beg : 0);
    if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
    // The band is walked in runs of up to eight columns that share one word of query codes: the word is fetched once per
    // run (and the next run's while this one computes), the codes come off it by shifting, and the cell of the next
    // column is read before this column's is written back, so neither load sits on the dependent path h -> f -> h.
    j = beg;
    uint32_t qnext = beg < end ? M.q[(size_t)(beg >> 3) * st] : 0;
    uint32_t p = beg < end ? M.eh[(size_t)beg * st] : 0;
    while (j < end) {
      uint32_t qs = qnext >> ((j & 7) << 2);
      const int jrun = (j | 7) + 1 < end ? (j | 7) + 1 : end;
      if (jrun < end) qnext = M.q[(size_t)(jrun >> 3) * st];
      for (; j < jrun; ++j, qs >>= 4) {
        const uint32_t pn = M.eh[(size_t)(j + 1) * st];   // (j + 1 <= end <= qlen: inside the array)
        const int q = (int)(qs & 15);
        const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
        int Mv = (int)(p & HMASK), e = (int)(p >> ESH);
        Mv = Mv ? Mv + s : 0;
        int h = Mv > e ? Mv : e;
        h = h > f ? h : f;
        const int hprev = h1;
        h1 = h;
        mj = m > h ? mj : j;
        m = m > h ? m : h;
        int t2 = Mv - oe_del; t2 = t2 > 0 ? t2 : 0;
        e -= e_del; e = e > t2 ? e : t2;
        M.eh[(size_t)j * st] = (Cell)((uint32_t)hprev | (uint32_t)e << ESH);
        t2 = Mv - oe_ins; t2 = t2 > 0 ? t2 : 0;
        f -= e_ins; f = f > t2 ? f : t2;
        p = pn;
      }
    }
    M.eh[(size_t)end * st] = (Cell)h1;  // {h1, e = 0}
    if (j == qlen) {
      max_ie = gscore > h1 ? max_ie : i;
      gscore = gscore > h1 ? gscore : h1;
    }
    if (m == 0) break;
    if (m > max) {
      max = m; max_i = i; max_j = mj;
      int off = mj - i; off = off < 0 ? -off : off;
      max_off = max_off > off ? max_off : off;
    } else if (zdrop > 0) {
      if (i - max_i > mj - max_j) {
        if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
      } else {
        if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
      }
    }
    for (j = beg; j < end && M.eh[(size_t)j * st] == 0; ++j) {}
    beg = j;
    for (j = end; j >= beg && M.eh[(size_t)j * st] == 0; --j) {}
    end = j + 2 < qlen ? j + 2 : qlen;
  }
  if (cells) *cells += ncell;
  *_qle = max_j + 1; *_tle = max_i + 1; *_gtle = max_ie + 1; *_gscore = gscore; *_max_off = max_off;
  return max;
}

// stage `len` codes produced by get(k) as 4-bit nibbles, 8 per word
template <class F>
B2_HD inline void b2_xext_stage(uint32_t* dst, int stride, int len, F get) {
  for (int k0 = 0; k0 < len; k0 += 8) {
    uint32_t wv = 0;
    for (int k = 0; k < 8 && k0 + k < len; ++k) wv |= (uint32_t)get(k0 + k) << (k << 2);
    dst[(size_t)(k0 >> 3) * stride] = wv;
  }
}
