// Group-cooperative (sub-warp) forms of the DP kernels, device only.
//
// A "task group" of B2_TASK_LANES lanes serves one read (extension stage) or one read pair
// (finalisation stage).  All lanes run the same control flow on the same data; inside the DP
// routines below the lanes split the cells of a row and exchange partial results with warp
// shuffles, so the per-read critical path shrinks by about the group width while the results stay
// bit-identical to the scalar forms in b2_ksw.cuh (which the host checker build keeps using):
//
//   b2_extend2_coop   row-parallel ksw_extend2 (bwa/ksw.c:380-479).  Within a row E depends only on
//                     the previous row and F is a max-plus prefix scan of M:
//                       F[j] = max_{beg<=k<j}( max(0, M[k]-oe_ins) + k*e_ins ) - (j-1)*e_ins
//                     each lane owns a contiguous chunk of the band, one exclusive prefix-max scan
//                     over the lanes links the chunks, and the row reductions (max with last-index
//                     tie rule, first/last non-zero cell for the band trimming) are butterflies.
//   b2_sw_local_coop  ksw_u8/ksw_i16 (bwa/ksw.c:111-334) with one GPU lane per SSE lane: the striped
//                     layout, the byte/word saturation, the lane shifts (_mm_slli_si128) and the
//                     lazy-F loop with its early exit map one to one.
#pragma once
#include "b2_ksw.cuh"
#include "b2_coopreg.cuh"

#define B2_TASK_LANES 16

#if defined(__CUDACC__)

struct B2G {
  unsigned mask;  // lanes of this group inside the warp
  int lane;       // 0..size-1
  int size;       // power of two
  __device__ inline int shfl(int v, int src) const { return __shfl_sync(mask, v, src, size); }
  __device__ inline int shfl_up(int v, int d) const { return __shfl_up_sync(mask, v, d, size); }
  __device__ inline int shfl_down(int v, int d) const { return __shfl_down_sync(mask, v, d, size); }
  __device__ inline int shfl_xor(int v, int m) const { return __shfl_xor_sync(mask, v, m, size); }
  __device__ inline long long shfl_xor64(long long v, int m) const { return __shfl_xor_sync(mask, v, m, size); }
  __device__ inline void sync() const { __syncwarp(mask); }
  // redux.sync: one instruction per 32-bit reduction over the lanes named in the mask
  __device__ inline int max_all(int v) const { return __reduce_max_sync(mask, v); }
  __device__ inline int min_all(int v) const { return __reduce_min_sync(mask, v); }
  // maximum of (hi, lo) pairs in lexicographic order, lo >= 0
  __device__ inline long long max_all64(long long v) const {
    const int hi = (int)(v >> 32);
    const int mh = __reduce_max_sync(mask, hi);
    const int lo = hi == mh ? (int)(v & 0x7fffffff) : -1;
    const int ml = __reduce_max_sync(mask, lo);
    return ((long long)mh << 32) | (unsigned)ml;
  }
};

// The same interface with the member mask known at compile time (low or high half of the warp): a shuffle with a
// run-time mask costs five extra instructions (MATCH.ANY / REDUX / VOTE / LOP3 / BRA.DIV convergence check).
template <unsigned MASK, int SIZE>
struct B2GC {
  int lane;
  static constexpr int size = SIZE;
  static constexpr unsigned mask = MASK;
  __device__ inline int shfl(int v, int src) const { return __shfl_sync(MASK, v, src, SIZE); }
  __device__ inline int shfl_up(int v, int d) const { return __shfl_up_sync(MASK, v, d, SIZE); }
  __device__ inline int shfl_down(int v, int d) const { return __shfl_down_sync(MASK, v, d, SIZE); }
  __device__ inline void sync() const { __syncwarp(MASK); }
  __device__ inline int max_all(int v) const { return __reduce_max_sync(MASK, v); }
  __device__ inline int min_all(int v) const { return __reduce_min_sync(MASK, v); }
};

B2_NOINLINE __device__ inline int b2_extend2_coop(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                      const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus,
                                      int zdrop, int h0, int* _qle, int* _tle, int* _gtle, int* _gscore, int* _max_off,
                                      B2EH* eh, unsigned long long* cells) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int G = g.size;
  int i, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
  {  // first row (closed form of the reference's serial initialisation)
    const int a1 = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = g.lane; j <= qlen; j += G) {
      int h = 0;
      if (j == 0) h = h0;
      else { h = a1 - (j - 1) * e_ins; h = h > 0 ? h : 0; }
      eh[j].h = h; eh[j].e = 0;
    }
  }
  g.sync();
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
  const int NEG = -(1 << 29);
  for (i = 0; i < tlen; ++i) {
    const int8_t* q = mat + target[i] * 5;
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
    if (cells && g.lane == 0) *cells += (unsigned long long)n;
    const int C = (n + G - 1) / G;
    const int mb = beg + g.lane * C;
    int me = mb + C; me = me < end ? me : end;
    // pass A: chunk-local maximum of V[k] = max(0, M[k]-oe_ins) + k*e_ins
    int vloc = NEG;
    for (int j = mb; j < me; ++j) {
      int M = eh[j].h;
      M = M ? M + q[query[j]] : 0;
      int t = M - oe_ins; t = t > 0 ? t : 0;
      t += j * e_ins;
      vloc = vloc > t ? vloc : t;
    }
    // exclusive prefix maximum over the lanes
    int incl = vloc;
    for (int d = 1; d < G; d <<= 1) { int o = g.shfl_up(incl, d); if (g.lane >= d) incl = incl > o ? incl : o; }
    int pre = g.shfl_up(incl, 1);
    if (g.lane == 0) pre = NEG;
    // pass B: cells of this lane, left to right
    long long key = -1;            // (H << 32 | j): row maximum, ties to the larger j
    int first_nz = 1 << 30, last_nz = -1;
    int lastH = 0, prevH = 0, pend_e = 0;
    for (int j = mb; j < me; ++j) {
      int M = eh[j].h, e = eh[j].e;
      M = M ? M + q[query[j]] : 0;
      int f = (j == beg) ? 0 : pre - (j - 1) * e_ins;
      int h = M > e ? M : e;
      h = h > f ? h : f;
      long long k = ((long long)h << 32) | (unsigned)j;
      key = key > k ? key : k;
      int t = M - oe_del; t = t > 0 ? t : 0;
      e -= e_del; e = e > t ? e : t;
      eh[j].e = e;
      if (j == mb) pend_e = e;  // its .h needs the left neighbour lane's last H
      else {
        eh[j].h = prevH;
        if (prevH | e) { first_nz = first_nz < j ? first_nz : j; last_nz = j; }
      }
      t = M - oe_ins; t = t > 0 ? t : 0;
      t += j * e_ins;
      pre = pre > t ? pre : t;
      prevH = h;
      lastH = h;
    }
    int inH = g.shfl_up(lastH, 1);
    if (mb < me) {
      int hh = (mb == beg) ? h1init : inH;
      eh[mb].h = hh;
      if (hh | pend_e) { first_nz = first_nz < mb ? first_nz : mb; last_nz = last_nz > mb ? last_nz : mb; }
    }
    const int last_lane = (n - 1) / C;
    const int h1 = g.shfl(lastH, last_lane);
    if (g.lane == last_lane) { eh[end].h = h1; eh[end].e = 0; }
    if (h1) last_nz = end;  // eh[end] = {h1, 0}
    key = g.max_all64(key);
    first_nz = g.min_all(first_nz);
    last_nz = g.max_all(last_nz);
    g.sync();
    const int m = (int)(key >> 32), mj = (int)(key & 0xffffffff);
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
    // band trimming: first non-zero cell of the new row, last non-zero cell (eh[end] included) + 2
    int nb = first_nz < end ? first_nz : end;          // no non-zero cell in [beg,end): beg = end
    int jl = last_nz >= nb ? last_nz : nb - 1;          // scanning down from `end` stops at nb-1
    beg = nb;
    end = jl + 2 < qlen ? jl + 2 : qlen;
  }
  if (_qle) *_qle = max_j + 1;
  if (_tle) *_tle = max_i + 1;
  if (_gtle) *_gtle = max_ie + 1;
  if (_gscore) *_gscore = gscore;
  if (_max_off) *_max_off = max_off;
  return max;
}

// Row-parallel ksw_global2 (bwa/ksw.c:504-606): the band [i-w, i+w] of row i is split over the lanes; F is the
// max-plus scan  F[j] = max( MINUS_INF-(j-beg)*e_ins, max_{beg<=k<j}( M[k]-oe_ins-(j-1-k)*e_ins ) ), which is the
// unrolled form of the reference's serial recurrence, so H, E, F and every direction bit (ties: m>=e, h>=f, e>t,
// f>t) are the same integers.  The traceback (a few hundred dependent steps) stays serial.
B2_NOINLINE __device__ inline int b2_global2_coop(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                      const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, B2EH* eh,
                                      uint8_t* z, int* n_cigar_, uint32_t* cigar, int cigar_cap, unsigned long long* cells) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int G = g.size;
  const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
  if (n_cigar_) *n_cigar_ = 0;
  for (int j = g.lane; j <= qlen; j += G) {
    if (j == 0) { eh[0].h = 0; eh[0].e = B2_MINUS_INF; }
    else if (j <= w) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = B2_MINUS_INF; }
    else { eh[j].h = B2_MINUS_INF; eh[j].e = B2_MINUS_INF; }
  }
  g.sync();
  const int NEG = -0x7f000000;  // below every value of the recurrence (MINUS_INF = -2^30 minus penalties)
  for (int i = 0; i < tlen; ++i) {
    const int8_t* q = mat + target[i] * 5;
    const int beg = i > w ? i - w : 0;
    const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
    const int h1init = beg == 0 ? -(o_del + e_del * (i + 1)) : B2_MINUS_INF;
    const int n = end - beg;
    if (n <= 0) { if (g.lane == 0) { eh[end].h = h1init; eh[end].e = B2_MINUS_INF; } g.sync(); continue; }
    if (cells && g.lane == 0) *cells += (unsigned long long)n;
    const int C = (n + G - 1) / G;
    const int mb = beg + g.lane * C;
    int me = mb + C; me = me < end ? me : end;
    int vloc = NEG;  // max over own cells of M[k] - oe_ins + k*e_ins
    for (int j = mb; j < me; ++j) {
      int v = (eh[j].h + q[query[j]]) - oe_ins + j * e_ins;
      vloc = vloc > v ? vloc : v;
    }
    int incl = vloc;
    for (int d = 1; d < G; d <<= 1) { int o = g.shfl_up(incl, d); if (g.lane >= d) incl = incl > o ? incl : o; }
    int pre = g.shfl_up(incl, 1);
    if (g.lane == 0) pre = NEG;
    uint8_t* zi = z ? z + (long)i * n_col : 0;
    int lastH = 0, prevH = 0;
    for (int j = mb; j < me; ++j) {
      int m = eh[j].h + q[query[j]], e = eh[j].e;
      const int f0 = B2_MINUS_INF - (j - beg) * e_ins;  // the serial recurrence's start value
      const int fs = pre == NEG ? NEG : pre - (j - 1) * e_ins;
      int f = j == beg ? B2_MINUS_INF : (fs > f0 ? fs : f0);
      uint8_t d = m >= e ? 0 : 1;
      int h = m >= e ? m : e;
      d = h >= f ? d : 2;
      h = h >= f ? h : f;
      int t = m - oe_del;
      e -= e_del;
      d |= e > t ? 1 << 2 : 0;
      e = e > t ? e : t;
      eh[j].e = e;
      t = m - oe_ins;
      d |= (f - e_ins) > t ? 2 << 4 : 0;
      if (zi) zi[j - beg] = d;
      if (j != mb) eh[j].h = prevH;
      const int v = m - oe_ins + j * e_ins;
      pre = pre > v ? pre : v;
      prevH = h; lastH = h;
    }
    int inH = g.shfl_up(lastH, 1);
    if (mb < me) eh[mb].h = (mb == beg) ? h1init : inH;
    const int last_lane = (n - 1) / C;
    const int h1 = g.shfl(lastH, last_lane);
    if (g.lane == last_lane) { eh[end].h = h1; eh[end].e = B2_MINUS_INF; }
    g.sync();
  }
  int score = eh[qlen].h;
  if (z && n_cigar_) {
    int n = 0, which = 0, k, i = tlen - 1;
    k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
      if (n + 2 > cigar_cap) { *n_cigar_ = -1; g.sync(); return score; }
      which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
      if (which == 0) { n = b2_push_cigar(cigar, n, 0, 1); --i; --k; }
      else if (which == 1) { n = b2_push_cigar(cigar, n, 2, 1); --i; }
      else { n = b2_push_cigar(cigar, n, 1, 1); --k; }
    }
    if (n + 2 > cigar_cap) { *n_cigar_ = -1; g.sync(); return score; }
    if (i >= 0) n = b2_push_cigar(cigar, n, 2, i + 1);
    if (k >= 0) n = b2_push_cigar(cigar, n, 1, k + 1);
    for (i = 0; i < n >> 1; ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
    *n_cigar_ = n;
  }
  g.sync();
  return score;
}

// Byte-mode striped SW with the whole DP state in registers: lane l owns query columns l*slen .. l*slen+slen-1
// (slen <= B2_SW_SLMAX, i.e. reads up to 16*B2_SW_SLMAX bases) as H/E/Hmax register arrays, and the substitution
// score comes from the base codes directly (the matrix is always bwa_fill_scmat's: a on the diagonal, -b off it,
// -1 against N; bwa/bwa.c:99-108), so the inner loops touch no memory at all.  Same lane-for-lane arithmetic as
// b2_sw_local_coop below (which remains the general path for word mode and longer reads).
#define B2_SW_SLMAX 16
// SL = columns per lane = ceil(qlen / 16), a compile-time constant: the register arrays have exactly that length, the
// column loops are fully unrolled without predication, and the substitution score of a column is one compare + select
// against per-column constants (the reference window comes from the 2-bit pac, so its codes are 0..3; rows with any
// other target code take the general expression).
template <int SL>
__device__ __noinline__ B2Kswr b2_sw_local_u8_regT(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                                   const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int xtra,
                                                   uint64_t* bbuf, unsigned long long* cells) {
  const int slen = SL;
  const int sa = mat[0], sb = mat[1], sn = mat[4];  // match, mismatch, ambiguous
  int shift = 127, mx = 0;
  for (int a = 0; a < 25; ++a) {
    if (mat[a] < (int8_t)shift) shift = mat[a];
    if (mat[a] > (int8_t)mx) mx = mat[a];
  }
  shift = (256 - shift) & 0xff;
  const int oe_d = o_del + e_del, oe_i = o_ins + e_ins;
  const bool use_scan = o_ins >= 1 && e_ins >= 1 && !(xtra & B2_KSW_XLITERAL);
  B2Kswr r; r.score = 0; r.te = r.qe = r.score2 = r.te2 = r.tb = r.qb = -1;
  const int minsc = (xtra & B2_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
  const int endsc = (xtra & B2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
  int n_b = 0, te = -1, gmax = 0;
  const int l = g.lane;
  int H[SL], E[SL], Hm[SL], qc[SL], sms[SL];   // H/E/saved row, base code, (mismatch-or-fixed score) + shift of the column
#pragma unroll
  for (int j = 0; j < SL; ++j) {
    H[j] = E[j] = Hm[j] = 0;
    const int col = l * slen + j;
    qc[j] = col < qlen ? query[col] : 5;  // 5: pad column, score 0
    sms[j] = shift + (qc[j] == 5 ? 0 : (qc[j] >= 4 ? sn : sb));
  }
  const int sas = sa + shift, sns = sn + shift;
  int last_b_row = -2, last_b_sc = 0;  // mirror of bbuf[n_b-1] kept in registers
  for (int i = 0; i < tlen; ++i) {
    const int tb = target[i];
    int f = 0, mxv = 0;
    int h = g.shfl_up(H[SL - 1], 1);   // H[slen-1] of the left lane (the byte shift of the last striped vector)
    if (l == 0) h = 0;
    if (tb < 4) {
#pragma unroll
      for (int j = 0; j < SL; ++j) {
        const int sp = qc[j] == tb ? sas : sms[j];
        int e = E[j], t;
        h += sp; h = h > 255 ? 255 : h;
        h -= shift; h = h < 0 ? 0 : h;
        h = h > e ? h : e;
        h = h > f ? h : f;
        mxv = mxv > h ? mxv : h;
        const int hn = H[j];
        H[j] = h;
        e -= e_del; e = e < 0 ? 0 : e;
        t = h - oe_d; t = t < 0 ? 0 : t;
        E[j] = e > t ? e : t;
        f -= e_ins; f = f < 0 ? 0 : f;
        t = h - oe_i; t = t < 0 ? 0 : t;
        f = f > t ? f : t;
        h = hn;
      }
    } else {
#pragma unroll
      for (int j = 0; j < SL; ++j) {
        const int sp = qc[j] == 5 ? shift : sns;
        int e = E[j], t;
        h += sp; h = h > 255 ? 255 : h;
        h -= shift; h = h < 0 ? 0 : h;
        h = h > e ? h : e;
        h = h > f ? h : f;
        mxv = mxv > h ? mxv : h;
        const int hn = H[j];
        H[j] = h;
        e -= e_del; e = e < 0 ? 0 : e;
        t = h - oe_d; t = t < 0 ? 0 : t;
        E[j] = e > t ? e : t;
        f -= e_ins; f = f < 0 ? 0 : f;
        t = h - oe_i; t = t < 0 ? 0 : t;
        f = f > t ? f : t;
        h = hn;
      }
    }
    if (use_scan) {  // carry of F across the lanes' column blocks (see b2_sw_local), then decay inside the block
      int x = f;
      const int D = slen * e_ins;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        int o = g.shfl_up(x, d) - d * D;
        if (l >= d && o > x) x = o;
      }
      f = g.shfl_up(x, 1);
      if (l == 0) f = 0;
#pragma unroll
      for (int j = 0; j < SL; ++j) {
        H[j] = H[j] > f ? H[j] : f;
        f -= e_ins; f = f < 0 ? 0 : f;
      }
    } else
    for (int k = 0, done = 0; k < 16 && !done; ++k) {  // lazy-F, literally
      f = g.shfl_up(f, 1);
      if (l == 0) f = 0;
#pragma unroll
      for (int j = 0; j < SL; ++j) {
        if (!done) {
          int hh = H[j];
          hh = hh > f ? hh : f;
          H[j] = hh;
          hh = hh - oe_i; if (hh < 0) hh = 0;
          f = f - e_ins; if (f < 0) f = 0;
          if (__all_sync(g.mask, !(f > hh))) done = 1;
        }
      }
    }
    const int imax = g.max_all(mxv);
    if (imax >= minsc) {
      if (n_b == 0 || last_b_row + 1 != i) { if (l == 0) bbuf[n_b] = (uint64_t)imax << 32 | (uint32_t)i; ++n_b; last_b_row = i; last_b_sc = imax; }
      else if (last_b_sc < imax) { if (l == 0) bbuf[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i; last_b_row = i; last_b_sc = imax; }
    }
    if (imax > gmax) {
      gmax = imax; te = i;
#pragma unroll
      for (int j = 0; j < SL; ++j) Hm[j] = H[j];
      if (gmax + shift >= 255 || gmax >= endsc) break;
    }
  }
  if (cells && l == 0) *cells += (unsigned long long)qlen * (unsigned long long)(te >= 0 && (gmax + shift >= 255 || gmax >= endsc) ? te + 1 : tlen);
  g.sync();
  r.score = gmax + shift < 255 ? gmax : 255;
  r.te = te;
  if (r.score != 255) {
    // qe: smallest query column holding the maximum of the saved row
    int best = -1, bcol = 1 << 30;
#pragma unroll
    for (int j = 0; j < SL; ++j) { int v = Hm[j] & 0xff, col = l * slen + j; if (v > best) { best = v; bcol = col; } }
    // reduce (max value, then min column)
    long long key = ((long long)best << 32) | (unsigned)(0x7fffffff - bcol);
    key = g.max_all64(key);
    r.qe = 0x7fffffff - (int)(key & 0xffffffff);
    if (n_b) {
      int ii = (r.score + mx - 1) / mx;
      int low = te - ii, high = te + ii;
      for (int i = 0; i < n_b; ++i) {
        int e = (int32_t)bbuf[i];
        if ((e < low || e > high) && (int)(bbuf[i] >> 32) > r.score2) { r.score2 = (int)(bbuf[i] >> 32); r.te2 = e; }
      }
    }
  }
  g.sync();
  return r;
}

__device__ inline B2Kswr b2_sw_local_u8_reg(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                            const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int xtra,
                                            uint64_t* bbuf, unsigned long long* cells) {
#define B2_SWR(N) case N: return b2_sw_local_u8_regT<N>(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, bbuf, cells);
  switch ((qlen + 15) / 16) {
    B2_SWR(1) B2_SWR(2) B2_SWR(3) B2_SWR(4) B2_SWR(5) B2_SWR(6) B2_SWR(7) B2_SWR(8)
    B2_SWR(9) B2_SWR(10) B2_SWR(11) B2_SWR(12) B2_SWR(13) B2_SWR(14) B2_SWR(15)
    default: return b2_sw_local_u8_regT<16>(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, bbuf, cells);
  }
#undef B2_SWR
}

// One GPU lane per SSE lane.  Lanes >= `lanes` of the group idle (16-bit mode uses 8 of 16).
__device__ inline B2Kswr b2_sw_local_coop(const B2G& g, int size, int qlen, const uint8_t* query, int tlen,
                                          const uint8_t* target, const int8_t* mat, int o_del, int e_del, int o_ins,
                                          int e_ins, int xtra, int16_t* work, uint64_t* bbuf, unsigned long long* cells) {
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
  int16_t *H0 = work, *H1 = work + n, *E = work + 2 * n, *Hmax = work + 3 * n;
  B2Kswr r; r.score = 0; r.te = r.qe = r.score2 = r.te2 = r.tb = r.qb = -1;
  const int minsc = (xtra & B2_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
  const int endsc = (xtra & B2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
  int n_b = 0, te = -1, gmax = 0;
  const int l = g.lane;
  const bool act = l < lanes;
  // sub-group of the active lanes (a contiguous, aligned block of `lanes` lanes inside the group)
  const unsigned amask = lanes == g.size ? g.mask : (g.mask & (((1u << lanes) - 1) << (__ffs(g.mask) - 1)));
  for (int i = l; i < n; i += g.size) { H0[i] = 0; E[i] = 0; Hmax[i] = 0; }
  g.sync();
  int broke = 0;
  for (int i = 0; i < tlen && !broke; ++i) {
    int imax = 0;
    if (act) {
      const int8_t* srow = mat + target[i] * 5;
      int f = 0, mxv = 0;
      int h = H0[(slen - 1) * lanes + l];
      h = __shfl_up_sync(amask, h, 1, lanes);
      if (l == 0) h = 0;
      for (int j = 0; j < slen; ++j) {
        const int col = l * slen + j;
        const int s = col < qlen ? srow[query[col]] : 0;
        int e = E[j * lanes + l], t;
        if (size == 1) {
          h = h + (s + shift); if (h > 255) h = 255;
          h = h - shift; if (h < 0) h = 0;
        } else {
          h = h + s; if (h > 32767) h = 32767; if (h < -32768) h = -32768;
        }
        h = h > e ? h : e;
        h = h > f ? h : f;
        mxv = mxv > h ? mxv : h;
        H1[j * lanes + l] = (int16_t)h;
        e = e - e_del; if (e < 0) e = 0;
        t = h - oe_d; if (t < 0) t = 0;
        e = e > t ? e : t;
        E[j * lanes + l] = (int16_t)e;
        f = f - e_ins; if (f < 0) f = 0;
        t = h - oe_i; if (t < 0) t = 0;
        f = f > t ? f : t;
        h = H0[j * lanes + l];
      }
      if (o_ins >= 1 && e_ins >= 1 && !(xtra & B2_KSW_XLITERAL)) {  // carry scan instead of the lazy-F loop (see b2_sw_local)
        int x = f;
        const int D = slen * e_ins;
        for (int d = 1; d < lanes; d <<= 1) {
          int o = __shfl_up_sync(amask, x, d, lanes) - d * D;
          if (l >= d && o > x) x = o;
        }
        f = __shfl_up_sync(amask, x, 1, lanes);
        if (l == 0) f = 0;
        for (int j = 0; j < slen && f > 0; ++j) {
          int hh = H1[j * lanes + l];
          H1[j * lanes + l] = (int16_t)(hh > f ? hh : f);
          f = f - e_ins;
        }
      } else
      // lazy-F
      for (int k = 0, done = 0; k < 16 && !done; ++k) {
        f = __shfl_up_sync(amask, f, 1, lanes);
        if (l == 0) f = 0;
        for (int j = 0; j < slen; ++j) {
          int hh = H1[j * lanes + l];
          hh = hh > f ? hh : f;
          H1[j * lanes + l] = (int16_t)hh;
          hh = hh - oe_i; if (hh < 0) hh = 0;
          f = f - e_ins; if (f < 0) f = 0;
          if (__all_sync(amask, !(f > hh))) { done = 1; break; }
        }
      }
      imax = mxv;
      for (int d = lanes >> 1; d; d >>= 1) { int o = __shfl_xor_sync(amask, imax, d, lanes); imax = imax > o ? imax : o; }
    }
    imax = g.shfl(imax, 0);
    if (cells && l == 0) *cells += (unsigned long long)qlen;
    if (imax >= minsc) {
      if (n_b == 0 || (int32_t)bbuf[n_b - 1] + 1 != i) { if (l == 0) bbuf[n_b] = (uint64_t)imax << 32 | (uint32_t)i; ++n_b; }
      else if ((int)(bbuf[n_b - 1] >> 32) < imax) { if (l == 0) bbuf[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i; }
      g.sync();
    }
    if (imax > gmax) {
      gmax = imax; te = i;
      if (act) for (int j = 0; j < slen; ++j) Hmax[j * lanes + l] = H1[j * lanes + l];
      if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) broke = 1;
    }
    g.sync();
    { int16_t* t = H1; H1 = H0; H0 = t; }
  }
  g.sync();
  if (size == 1) r.score = gmax + shift < 255 ? gmax : 255; else r.score = gmax;
  r.te = te;
  if (size != 1 || r.score != 255) {
    int max = -1;
    if (size != 1) r.qe = -1;
    for (int i = 0; i < n; ++i) {
      int v = (size == 1) ? (Hmax[i] & 0xff) : (uint16_t)Hmax[i];
      int col = i / lanes + i % lanes * slen;
      if (v > max) { max = v; r.qe = col; }
      else if (v == max && col < r.qe) r.qe = col;
    }
    if (n_b) {
      int ii = (r.score + mx - 1) / mx;
      int low = te - ii, high = te + ii;
      for (int i = 0; i < n_b; ++i) {
        int e = (int32_t)bbuf[i];
        if ((e < low || e > high) && (int)(bbuf[i] >> 32) > r.score2) { r.score2 = (int)(bbuf[i] >> 32); r.te2 = e; }
      }
    }
  }
  g.sync();
  return r;
}

// ---- dispatch to the register-resident forms (b2_coopreg.cuh) -------------------------------------------------
// Whole-warp groups only (the member mask is then the constant 0xffffffff: a shuffle with a run-time mask costs
// five extra instructions -- MATCH.ANY / REDUX / VOTE / LOP3 / BRA.DIV -- for the convergence check).  One
// out-of-line instance per cells-per-lane count, shared by the call sites (mem_chain2aln's two extensions, the
// several users of bwa_gen_cigar2).  Returns false when the preconditions do not hold (the callers then use the
// memory-resident forms above).
typedef B2GC<0xffffffffu, 32> B2GW;

template <int C>
__device__ __noinline__ int b2_extend2_reg_dev(int lane, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                               const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus,
                                               int zdrop, int h0, int* res5, unsigned long long* cells) {
  B2GW g; g.lane = lane;
  return b2_extend2_reg<C>(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, res5, res5 + 1,
                           res5 + 2, res5 + 3, res5 + 4, cells);
}

__device__ inline bool b2_extend2_reg_try(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                          const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus,
                                          int zdrop, int h0, int* qle, int* tle, int* gtle, int* gscore, int* max_off,
                                          unsigned long long* cells, int* score) {
  if (g.size != 32 || qlen + 1 > 32 * 8 || !b2_mat_simple(mat)) return false;
  const int need = (qlen + 1 + 31) >> 5;
  int r[5];
  int s;
#define B2_EXT_ARGS g.lane, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r, cells
  if (need <= 1) s = b2_extend2_reg_dev<1>(B2_EXT_ARGS);
  else if (need <= 2) s = b2_extend2_reg_dev<2>(B2_EXT_ARGS);
  else if (need <= 3) s = b2_extend2_reg_dev<3>(B2_EXT_ARGS);
  else if (need <= 5) s = b2_extend2_reg_dev<5>(B2_EXT_ARGS);
  else s = b2_extend2_reg_dev<8>(B2_EXT_ARGS);
#undef B2_EXT_ARGS
  if (qle) *qle = r[0];
  if (tle) *tle = r[1];
  if (gtle) *gtle = r[2];
  if (gscore) *gscore = r[3];
  if (max_off) *max_off = r[4];
  *score = s;
  return true;
}

template <int C>
__device__ __noinline__ int b2_global2_diag_dev(int lane, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                                const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, uint8_t* z,
                                                int* n_cigar_, uint32_t* cigar, int cigar_cap, unsigned long long* cells) {
  B2GW g; g.lane = lane;
  return b2_global2_diag<C>(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, z, n_cigar_, cigar, cigar_cap, cells);
}

__device__ inline bool b2_global2_diag_try(const B2G& g, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                           const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, uint8_t* z,
                                           int* n_cigar_, uint32_t* cigar, int cigar_cap, unsigned long long* cells, int* score) {
  const int dl = qlen > tlen ? qlen - tlen : tlen - qlen;
  if (g.size != 32 || qlen < 1 || tlen < 1 || dl > w || 2 * w + 1 > 32 * 4 || !b2_mat_simple(mat)) return false;
  const int need = (2 * w + 1 + 31) >> 5;
#define B2_GLB_ARGS g.lane, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, z, n_cigar_, cigar, cigar_cap, cells
  if (need <= 1) *score = b2_global2_diag_dev<1>(B2_GLB_ARGS);
  else if (need <= 2) *score = b2_global2_diag_dev<2>(B2_GLB_ARGS);
  else *score = b2_global2_diag_dev<4>(B2_GLB_ARGS);
#undef B2_GLB_ARGS
  return true;
}

#endif  // __CUDACC__
