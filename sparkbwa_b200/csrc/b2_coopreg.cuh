// Register-resident group-cooperative DP (no memory traffic inside a row).  The routines are templates over
// the lane-group type: the device passes B2G (warp shuffles, b2_coop.cuh), tests/coop_fuzz.cpp passes a
// coroutine emulation of the same interface so that they are fuzzed against the scalar forms of b2_ksw.cuh
// on a machine without a GPU.
//
//   b2_global2_diag<C>  ksw_global2 (bwa/ksw.c:504-606) with the band stored by DIAGONAL: lane l owns the
//                       diagonals d = j - i + w in [l*C, l*C + C).  The diagonal predecessor H[i-1][j-1] then
//                       lives in the same register of the same lane, E[i][j] (computed at (i-1, j), diagonal
//                       d+1) arrives with one neighbour shuffle, and F is the max-plus prefix scan of the row.
//                       Narrow bands (w <= 7: one cell per lane per row) are the common case of mem_reg2aln.
//   b2_extend2_reg<C>   ksw_extend2 (bwa/ksw.c:380-479) with the eh[] array of the reference held in registers,
//                       column-blocked: lane l owns eh[l*C .. l*C + C).  Same row-parallel restatement as
//                       b2_extend2_coop, without the loads/stores and with a fixed cell
//                       schedule that the compiler fully unrolls.
// Both need the substitution matrix to have the shape bwa_fill_scmat (bwa/bwa.c:99-108) produces -- a on the
// diagonal, -b off it, one value against N -- so the score is a compare/select on the base codes.
#pragma once
#include "b2_ksw.cuh"

#if defined(__CUDACC__)
#define B2_UNROLL _Pragma("unroll")
#else
#define B2_UNROLL
#endif

B2_HD inline bool b2_mat_simple(const int8_t* mat) {
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) {
      const int v = mat[i * 5 + j];
      const int want = (i == 4 || j == 4) ? mat[4] : (i == j ? mat[0] : mat[1]);
      if (v != want) return false;
    }
  return true;
}

// a[idx] of a register array without dynamic indexing
template <int C>
B2_HD inline int b2_sel(const int (&a)[C], int idx) {
  int v = 0;
B2_UNROLL
  for (int c = 0; c < C; ++c) if (c == idx) v = a[c];
  return v;
}

// traceback of ksw_global2 (bwa/ksw.c:585-604) over the direction matrix z (row stride n_col)
B2_HD inline void b2_global2_backtrack(const uint8_t* z, int n_col, int w, int qlen, int tlen, int* n_cigar_,
                                       uint32_t* cigar, int cigar_cap) {
  int n = 0, which = 0, k, i = tlen - 1;
  k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
  while (i >= 0 && k >= 0) {
    if (n + 2 > cigar_cap) { *n_cigar_ = -1; return; }
    which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
    if (which == 0) { n = b2_push_cigar(cigar, n, 0, 1); --i; --k; }
    else if (which == 1) { n = b2_push_cigar(cigar, n, 2, 1); --i; }
    else { n = b2_push_cigar(cigar, n, 1, 1); --k; }
  }
  if (n + 2 > cigar_cap) { *n_cigar_ = -1; return; }
  if (i >= 0) n = b2_push_cigar(cigar, n, 2, i + 1);
  if (k >= 0) n = b2_push_cigar(cigar, n, 1, k + 1);
  for (i = 0; i < n >> 1; ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
  *n_cigar_ = n;
}

// Preconditions (checked by the caller): b2_mat_simple(mat), qlen >= 1, tlen >= 1, |qlen - tlen| <= w,
// 2*w + 1 <= C * g.size, C <= 10.  Every lane returns the score; lane 0's `cells` is advanced.
template <int C, class G>
B2_HD inline int b2_global2_diag(const G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                 const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, uint8_t* z,
                                 int* n_cigar_, uint32_t* cigar, int cigar_cap, unsigned long long* cells) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int GS = g.size, l = g.lane;
  const int nd = 2 * w + 1;
  const int n_col = qlen < nd ? qlen : nd;
  const int sa = mat[0], sb = mat[1], sn = mat[4];
  const int NEG = -0x7f000000;
  const int d0 = l * C;
  if (n_cigar_) *n_cigar_ = 0;
  int hp[C], en[C];
  uint32_t qpk = 0;  // 3-bit base codes of the C columns under this lane's diagonals (slides one column per row)
B2_UNROLL
  for (int c = 0; c < C; ++c) {
    const int j = d0 + c - w;  // column of this diagonal in row 0
    hp[c] = (j >= 1 && j <= w) ? -(o_ins + e_ins * j) : B2_MINUS_INF;
    en[c] = B2_MINUS_INF;
    qpk |= (uint32_t)((j >= 0 && j < qlen) ? query[j] : 4) << (3 * c);
  }
  unsigned long long ncell = 0;
  int tch = 0, last_h = 0;
  const int d_fin = qlen - tlen + w;  // diagonal of the cell (tlen-1, qlen-1)
  for (int i = 0; i < tlen; ++i) {
    if ((i & (GS - 1)) == 0) tch = (i + l < tlen) ? target[i + l] : 4;
    const int t = g.shfl(tch, i & (GS - 1));
    const int dmin = w > i ? w - i : 0;                          // diagonal of column beg
    const int dend = qlen - i + w < nd ? qlen - i + w : nd;      // one past the diagonal of column end-1
    const int hb = i ? -(o_del + e_del * i) : 0;                 // H[i-1][-1]
    ncell += (unsigned long long)(dend - dmin);
    // E[i][j] was computed at (i-1, j): the next diagonal of the previous row
    int ein = g.shfl_down(en[0], 1);
    if (l == GS - 1) ein = B2_MINUS_INF;
    int vloc = NEG;
B2_UNROLL
    for (int c = 0; c < C; ++c) {
      const int d = d0 + c;
      const int q = (int)(qpk >> (3 * c)) & 7;
      const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
      const int hsrc = d == dmin && i <= w ? hb : hp[c];         // column 0 takes the boundary value
      const int v = hsrc + s - oe_ins + d * e_ins;
      if (d >= dmin && d < dend) vloc = vloc > v ? vloc : v;
    }
    int incl = vloc;
    for (int k = 1; k < GS; k <<= 1) { const int o = g.shfl_up(incl, k); if (l >= k) incl = incl > o ? incl : o; }
    int pre = g.shfl_up(incl, 1);
    if (l == 0) pre = NEG;
    uint8_t* zi = z ? z + (long)i * n_col - dmin : 0;
B2_UNROLL
    for (int c = 0; c < C; ++c) {
      const int d = d0 + c;
      if (d >= dmin && d < dend) {
        const int q = (int)(qpk >> (3 * c)) & 7;
        const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
        const int mm = (d == dmin && i <= w ? hb : hp[c]) + s;
        int e = c + 1 < C ? en[c + 1] : ein;                     // (en[c + 1] still holds the previous row)
        const int f0 = B2_MINUS_INF - (d - dmin) * e_ins;
        const int fs = pre == NEG ? NEG : pre - (d - 1) * e_ins;
        const int f = d == dmin ? B2_MINUS_INF : (fs > f0 ? fs : f0);
        uint8_t dir = mm >= e ? 0 : 1;
        int h = mm >= e ? mm : e;
        dir = h >= f ? dir : 2;
        h = h >= f ? h : f;
        int t2 = mm - oe_del;
        e -= e_del;
        dir |= e > t2 ? 1 << 2 : 0;
        e = e > t2 ? e : t2;
        t2 = mm - oe_ins;
        dir |= (f - e_ins) > t2 ? 2 << 4 : 0;
        if (zi) zi[d] = dir;
        hp[c] = h; en[c] = e;
        const int v = mm - oe_ins + d * e_ins;
        pre = pre > v ? pre : v;
        if (d == d_fin) last_h = h;
      }
    }
    // slide the query window: diagonal d is at column j+1 in the next row
    int qin = g.shfl_down((int)(qpk & 7), 1);
    if (l == GS - 1) { const int j = i + 1 - w + GS * C - 1; qin = (j >= 0 && j < qlen) ? query[j] : 4; }
    qpk = (qpk >> 3) | ((uint32_t)qin << (3 * (C - 1)));
  }
  if (cells && l == 0) *cells += ncell;
  const int score = g.shfl(last_h, d_fin / C);
  if (z && n_cigar_) {
    g.sync();
    b2_global2_backtrack(z, n_col, w, qlen, tlen, n_cigar_, cigar, cigar_cap);
  }
  g.sync();
  return score;
}

// Preconditions: b2_mat_simple(mat), qlen + 1 <= C * g.size, C <= 20.
template <int C, class G>
B2_HD inline int b2_extend2_reg(const G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus,
                                int zdrop, int h0, int* _qle, int* _tle, int* _gtle, int* _gscore, int* _max_off,
                                unsigned long long* cells) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int GS = g.size, l = g.lane;
  const int sa = mat[0], sb = mat[1], sn = mat[4];
  const int NEG = -(1 << 29);
  const int j0 = l * C;
  int H[C], E[C];               // eh[j0 + c].h, eh[j0 + c].e of the reference
  uint32_t qpk[(C + 9) / 10];   // query[j0 + c] as 3-bit codes
  {
    const int a1 = h0 > oe_ins ? h0 - oe_ins : 0;
B2_UNROLL
    for (int k = 0; k < (C + 9) / 10; ++k) qpk[k] = 0;
B2_UNROLL
    for (int c = 0; c < C; ++c) {
      const int j = j0 + c;
      int h = 0;
      if (j == 0) h = h0;
      else if (j <= qlen) { h = a1 - (j - 1) * e_ins; h = h > 0 ? h : 0; }
      H[c] = h; E[c] = 0;
      qpk[c / 10] |= (uint32_t)(j < qlen ? query[j] : 4) << (3 * (c % 10));
    }
  }
  int i, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
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
  int tch = 0;
  for (i = 0; i < tlen; ++i) {
    if ((i & (GS - 1)) == 0) tch = (i + l < tlen) ? target[i + l] : 4;
    const int t = g.shfl(tch, i & (GS - 1));
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    const int n = end - beg;
    int h1init;
    if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; } else h1init = 0;
    if (n <= 0) {  // empty band: the reference's inner loop does not run, m stays 0
      if (end == qlen) { max_ie = gscore > h1init ? max_ie : i; gscore = gscore > h1init ? gscore : h1init; }
      break;
    }
    ncell += (unsigned long long)n;
    // pass A: chunk maximum of V[k] = max(0, M[k]-oe_ins) + k*e_ins over the own cells of the band
    int vloc = NEG;
B2_UNROLL
    for (int c = 0; c < C; ++c) {
      const int j = j0 + c;
      const int q = (int)(qpk[c / 10] >> (3 * (c % 10))) & 7;
      const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
      int mm = H[c];
      mm = mm ? mm + s : 0;
      int v = mm - oe_ins; v = v > 0 ? v : 0;
      v += j * e_ins;
      if (j >= beg && j < end) vloc = vloc > v ? vloc : v;
    }
    int incl = vloc;
    for (int k = 1; k < GS; k <<= 1) { const int o = g.shfl_up(incl, k); if (l >= k) incl = incl > o ? incl : o; }
    int pre = g.shfl_up(incl, 1);
    if (l == 0) pre = NEG;
    // pass B: the cells left to right.  eh[j].h <- H[i][j-1] for beg < j <= end (eh[beg].h <- h1init, eh[end].e <- 0)
    // happens in place once the old value has been consumed; the first own column waits for the neighbour's H.
    int bh = -1, bj = -1;                    // row maximum of this lane, ties to the larger j
    int first_nz = 1 << 30, last_nz = -1;
    int hprev = 0, hend = 0;
B2_UNROLL
    for (int c = 0; c < C; ++c) {
      const int j = j0 + c;
      int h = 0;
      const bool inb = j >= beg && j < end;
      if (inb) {
        const int q = (int)(qpk[c / 10] >> (3 * (c % 10))) & 7;
        const int s = (q | t) >= 4 ? sn : (q == t ? sa : sb);
        int mm = H[c];
        mm = mm ? mm + s : 0;
        int e = E[c];
        const int f = (j == beg) ? 0 : pre - (j - 1) * e_ins;
        h = mm > e ? mm : e;
        h = h > f ? h : f;
        if (h >= bh) { bh = h; bj = j; }
        int t2 = mm - oe_del; t2 = t2 > 0 ? t2 : 0;
        e -= e_del; e = e > t2 ? e : t2;
        E[c] = e;
        t2 = mm - oe_ins; t2 = t2 > 0 ? t2 : 0;
        t2 += j * e_ins;
        pre = pre > t2 ? pre : t2;
        if (j == end - 1) hend = h;
      }
      if (c > 0 && j >= beg && j <= end) {
        const int hh = (j == beg) ? h1init : hprev;
        H[c] = hh;
        if (j == end) E[c] = 0;
        if (hh | E[c]) { if (j < end) first_nz = first_nz < j ? first_nz : j; last_nz = j; }
      }
      hprev = h;
    }
    {
      const int hin = g.shfl_up(hprev, 1);   // H[i][j0 - 1]: the left neighbour's last column
      if (j0 >= beg && j0 <= end) {
        const int hh = (j0 == beg) ? h1init : hin;
        H[0] = hh;
        if (j0 == end) E[0] = 0;
        if (hh | E[0]) { if (j0 < end) first_nz = first_nz < j0 ? first_nz : j0; last_nz = last_nz > j0 ? last_nz : j0; }
      }
    }
    const int h1 = g.shfl(hend, (end - 1) / C);  // H[i][end-1]
    const int m = g.max_all(bh);
    const int mj = g.max_all(bh == m ? bj : -1);
    first_nz = g.min_all(first_nz);
    last_nz = g.max_all(last_nz);
    if (end == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
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
    // band trimming: first non-zero cell of [beg, end), last non-zero cell of [beg, end] + 2
    // (the reference's first scan stops at end: when only eh[end] is non-zero, beg becomes end)
    const int nb = first_nz < end ? first_nz : end;
    const int jl = last_nz >= nb ? last_nz : nb - 1;
    beg = nb;
    end = jl + 2 < qlen ? jl + 2 : qlen;
  }
  if (cells && l == 0) *cells += ncell;
  if (_qle) *_qle = max_j + 1;
  if (_tle) *_tle = max_i + 1;
  if (_gtle) *_gtle = max_ie + 1;
  if (_gscore) *_gscore = gscore;
  if (_max_off) *_max_off = max_off;
  return max;
}

// ---------------------------------------------------------------------------------------------
// ksw_u8 / ksw_i16 (bwa/ksw.c:111-334) for a lane group of ANY size -- down to a single lane, i.e. fewer lanes than the
// reference's 16 (byte) / 8 (word) SSE lanes.  This is the model of SURVEY.md Appendix A in query-column order: the
// padded query (n = slen * lanes columns) is cut into `lanes` blocks of slen columns; block b is what SSE lane b of the
// reference holds.  Per target row: (1) every block runs left to right with its F chain restarted at the block start
// (first pass; the row maximum and the new E come from this pass only), (2) the F values leaving the blocks are
// combined into the carry each block receives from its left -- x_b = max(F_b, x_{b-1} - slen*e_ins), what the lazy-F
// loop converges to -- and decay into the block raising H (E is NOT revisited).  The blocks are dealt over the lanes
// of the group (block b -> lane b mod G); the only exchange is the 2 x lanes values of step (2), through memory.
// With o_ins or e_ins below 1 the reference's lazy-F early exit is observable: that case runs the
// literal loop on one lane.  work: 4*n + 32 int16 (H0, H1, E, Hmax, exchange); bbuf: tlen+1 entries.
template <class G>
B2_HD inline B2Kswr b2_sw_local_g(const G& g, int size, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                  const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int xtra, int16_t* work,
                                  uint64_t* bbuf, unsigned long long* cells) {
  const int lanes = size == 1 ? 16 : 8;
  const int slen = (qlen + lanes - 1) / lanes;
  const int n = slen * lanes;
  int shift = 127, mx = 0;
  for (int a = 0; a < 25; ++a) {
    if (mat[a] < (int8_t)shift) shift = mat[a];
    if (mat[a] > (int8_t)mx) mx = mat[a];
  }
  shift = (256 - shift) & 0xff;
  const int oe_d = o_del + e_del, oe_i = o_ins + e_ins;
  int16_t *H0 = work, *H1 = work + n, *E = work + 2 * n, *Hmax = work + 3 * n, *XF = work + 4 * n, *XM = work + 4 * n + 16;
  B2Kswr r; r.score = 0; r.te = r.qe = r.score2 = r.te2 = r.tb = r.qb = -1;
  const int minsc = (xtra & B2_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
  const int endsc = (xtra & B2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
  const bool scan = o_ins >= 1 && e_ins >= 1 && !(xtra & B2_KSW_XLITERAL);
  int n_b = 0, te = -1, gmax = 0, last_row = -2, last_sc = 0;
  for (int p = g.lane; p < n; p += g.size) { H0[p] = 0; E[p] = 0; Hmax[p] = 0; }
  g.sync();
  for (int i = 0; i < tlen; ++i) {
    const int8_t* srow = mat + target[i] * 5;
    for (int b = g.lane; b < lanes; b += g.size) {  // (1) first pass over the blocks this lane owns
      const int p0 = b * slen;
      int f = 0, mxv = 0;
      int h = b ? H0[p0 - 1] : 0;
      for (int p = p0; p < p0 + slen; ++p) {
        const int s = p < qlen ? srow[query[p]] : 0;
        int e = E[p], t;
        if (size == 1) { h += s + shift; h = h > 255 ? 255 : h; h -= shift; h = h < 0 ? 0 : h; }
        else { h += s; h = h > 32767 ? 32767 : h; h = h < -32768 ? -32768 : h; }
        h = h > e ? h : e;
        h = h > f ? h : f;
        mxv = mxv > h ? mxv : h;
        H1[p] = (int16_t)h;
        e -= e_del; e = e < 0 ? 0 : e;
        t = h - oe_d; t = t < 0 ? 0 : t;
        E[p] = (int16_t)(e > t ? e : t);
        f -= e_ins; f = f < 0 ? 0 : f;
        t = h - oe_i; t = t < 0 ? 0 : t;
        f = f > t ? f : t;
        h = H0[p];
      }
      XF[b] = (int16_t)f; XM[b] = (int16_t)mxv;
    }
    g.sync();
    int imax = 0;
    for (int b = 0; b < lanes; ++b) imax = imax > XM[b] ? imax : XM[b];
    if (scan) {  // (2) carry into each block, then decay inside it
      int x = 0;
      for (int b = 1; b < lanes; ++b) {
        int c = x - slen * e_ins; c = c < 0 ? 0 : c;
        x = XF[b - 1] > c ? XF[b - 1] : c;           // x = carry leaving block b-1
        if (b % g.size != g.lane) continue;
        int f = x;
        for (int p = b * slen; p < (b + 1) * slen && f > 0; ++p) {
          if (H1[p] < f) H1[p] = (int16_t)f;
          f -= e_ins;
        }
      }
    } else if (g.lane == 0) {  // the reference's lazy-F loop as it is (vector element b of step j = column b*slen + j)
      int fv[16];
      for (int b = 0; b < lanes; ++b) fv[b] = XF[b];
      for (int k = 0, done = 0; k < 16 && !done; ++k) {
        for (int b = lanes - 1; b > 0; --b) fv[b] = fv[b - 1];
        fv[0] = 0;
        for (int j = 0; j < slen; ++j) {
          int all = 1;
          for (int b = 0; b < lanes; ++b) {
            const int p = b * slen + j;
            int h = H1[p];
            h = h > fv[b] ? h : fv[b];
            H1[p] = (int16_t)h;
            h -= oe_i; h = h < 0 ? 0 : h;
            int f = fv[b] - e_ins; f = f < 0 ? 0 : f;
            fv[b] = f;
            if (f > h) all = 0;
          }
          if (all) { done = 1; break; }
        }
      }
    }
    if (cells && g.lane == 0) *cells += (unsigned long long)qlen;
    if (imax >= minsc) {  // sub-optimal list: consecutive rows merge into one entry (bwa/ksw.c:192-199)
      if (n_b == 0 || last_row + 1 != i) { if (g.lane == 0) bbuf[n_b] = (uint64_t)imax << 32 | (uint32_t)i; ++n_b; last_row = i; last_sc = imax; }
      else if (last_sc < imax) { if (g.lane == 0) bbuf[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i; last_row = i; last_sc = imax; }
    }
    g.sync();
    int stop = 0;
    if (imax > gmax) {
      gmax = imax; te = i;
      for (int p = g.lane; p < n; p += g.size) Hmax[p] = H1[p];
      if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) stop = 1;
    }
    if (stop) break;
    { int16_t* t = H1; H1 = H0; H0 = t; }
  }
  g.sync();
  r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
  r.te = te;
  if (size != 1 || r.score != 255) {
    int max = -1;
    if (size != 1) r.qe = -1;
    for (int p = 0; p < n; ++p) {  // smallest query column holding the maximum of the saved row
      const int v = size == 1 ? (Hmax[p] & 0xff) : (uint16_t)Hmax[p];
      if (v > max) { max = v; r.qe = p; }
    }
    if (n_b) {
      const int ii = (r.score + mx - 1) / mx;
      const int low = te - ii, high = te + ii;
      for (int k = 0; k < n_b; ++k) {
        const int e = (int32_t)bbuf[k];
        if ((e < low || e > high) && (int)(bbuf[k] >> 32) > r.score2) { r.score2 = (int)(bbuf[k] >> 32); r.te2 = e; }
      }
    }
  }
  g.sync();
  return r;
}

// ksw_align2 (bwa/ksw.c:343-365) on top of it: end points and second best, then the start by a rerun on the reversed
// prefixes (full tlen, KSW_XSTOP|score).  query/target are reversed in place and restored, by one lane.
template <class G>
B2_HD inline B2Kswr b2_align2_g(const G& g, int qlen, uint8_t* query, int tlen, uint8_t* target, const int8_t* mat, int o_del,
                                int e_del, int o_ins, int e_ins, int xtra, int16_t* work, uint64_t* bbuf, unsigned long long* cells) {
  const int size = (xtra & B2_KSW_XBYTE) ? 1 : 2;
  B2Kswr r = b2_sw_local_g(g, size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, work, bbuf, cells);
  if ((xtra & B2_KSW_XSTART) == 0 || ((xtra & B2_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
  if (g.lane == 0) { b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target); }
  g.sync();
  B2Kswr rr = b2_sw_local_g(g, size, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins,
                            B2_KSW_XSTOP | r.score | (xtra & B2_KSW_XLITERAL), work, bbuf, cells);
  if (g.lane == 0) { b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target); }
  g.sync();
  if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
  return r;
}
