// TEST INFRASTRUCTURE ONLY -- not compiled into libbwa.so.
// Plain one-alignment-at-a-time scalar forms of the DP routines, used by the host checker build (tests/hostsim, where
// a kernel launch is a serial loop) and as the reference side of the fuzzers (tests/coop_fuzz.cpp, tests/sw_fuzz.cpp).
// They follow the reference's loops closely on purpose -- that is what makes them a checker:
//   b2_extend2  <- ksw_extend2  bwa/ksw.c:380-479   banded affine-gap extension with z-drop
//   b2_global2  <- ksw_global2  bwa/ksw.c:504-606   banded global alignment + traceback
//   b2_sw_local <- ksw_u8 / ksw_i16 bwa/ksw.c:111-334  striped local SW, restated lane-for-lane
//   b2_align2   <- ksw_align2   bwa/ksw.c:343-365   end, second best, then start by reversed rerun
// Tie rules, band trimming and break conditions follow SURVEY.md H5-H7; every quantity is an integer.  The striped
// local kernel keeps the reference's 16 (byte) / 8 (word) SSE lanes as an explicit inner dimension, including the
// lazy-F loop and its early exit, because those artefacts are visible in (score, te, qe, score2).
// The shipped library computes the same results with the lane-group / one-alignment-per-lane kernels of
// sparkbwa_b200/csrc/b2_coop.cuh, b2_coopreg.cuh, b2_xext.cuh and b2_galn.cuh.
#pragma once

// query/target: codes 0..4; eh: qlen+1 entries of scratch.  Returns the best extension score.
B2_NOINLINE B2_HD inline int b2_extend2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, const int8_t* mat,
                            int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                            int* _qle, int* _tle, int* _gtle, int* _gscore, int* _max_off, B2EH* eh,
                            unsigned long long* cells = 0) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  int i, j, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
  for (j = 0; j <= qlen; ++j) eh[j].h = eh[j].e = 0;
  eh[0].h = h0;
  if (qlen >= 1) eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;  // (the reference writes eh[1] even for qlen == 0)
  for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;
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
  for (i = 0; i < tlen; ++i) {
    int t, f = 0, h1, m = 0, mj = -1;
    const int8_t* q = mat + target[i] * 5;
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    if (cells) *cells += (unsigned long long)(end > beg ? end - beg : 0);
    if (beg == 0) {
      h1 = h0 - (o_del + e_del * (i + 1));
      if (h1 < 0) h1 = 0;
    } else h1 = 0;
    for (j = beg; j < end; ++j) {
      B2EH* p = &eh[j];
      int h, M = p->h, e = p->e;
      p->h = h1;
      M = M ? M + q[query[j]] : 0;
      h = M > e ? M : e;
      h = h > f ? h : f;
      h1 = h;
      mj = m > h ? mj : j;
      m = m > h ? m : h;
      t = M - oe_del; t = t > 0 ? t : 0;
      e -= e_del; e = e > t ? e : t;
      p->e = e;
      t = M - oe_ins; t = t > 0 ? t : 0;
      f -= e_ins; f = f > t ? f : t;
    }
    eh[end].h = h1; eh[end].e = 0;
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
    for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) {}
    beg = j;
    for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) {}
    end = j + 2 < qlen ? j + 2 : qlen;
  }
  if (_qle) *_qle = max_j + 1;
  if (_tle) *_tle = max_i + 1;
  if (_gtle) *_gtle = max_ie + 1;
  if (_gscore) *_gscore = gscore;
  if (_max_off) *_max_off = max_off;
  return max;
}

// Banded global alignment.  z: backtrack matrix of min(qlen,2w+1)*tlen bytes, or null for score only.
// cigar (capacity cigar_cap) receives the operations; *n_cigar < 0 signals capacity overflow.
B2_NOINLINE B2_HD inline int b2_global2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, const int8_t* mat,
                            int o_del, int e_del, int o_ins, int e_ins, int w, B2EH* eh, uint8_t* z,
                            int* n_cigar_, uint32_t* cigar, int cigar_cap, unsigned long long* cells = 0) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  int i, j, n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
  if (n_cigar_) *n_cigar_ = 0;
  eh[0].h = 0; eh[0].e = B2_MINUS_INF;
  for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = B2_MINUS_INF; }
  for (; j <= qlen; ++j) eh[j].h = eh[j].e = B2_MINUS_INF;
  for (i = 0; i < tlen; ++i) {
    int32_t f = B2_MINUS_INF, h1, beg, end, t;
    const int8_t* q = mat + target[i] * 5;
    beg = i > w ? i - w : 0;
    end = i + w + 1 < qlen ? i + w + 1 : qlen;
    h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : B2_MINUS_INF;
    if (cells) *cells += (unsigned long long)(end > beg ? end - beg : 0);
    uint8_t* zi = z ? z + (long)i * n_col : 0;
    for (j = beg; j < end; ++j) {
      B2EH* p = &eh[j];
      int32_t h, m = p->h, e = p->e;
      uint8_t d;
      p->h = h1;
      m += q[query[j]];
      d = m >= e ? 0 : 1;
      h = m >= e ? m : e;
      d = h >= f ? d : 2;
      h = h >= f ? h : f;
      h1 = h;
      t = m - oe_del;
      e -= e_del;
      d |= e > t ? 1 << 2 : 0;
      e = e > t ? e : t;
      p->e = e;
      t = m - oe_ins;
      f -= e_ins;
      d |= f > t ? 2 << 4 : 0;
      f = f > t ? f : t;
      if (zi) zi[j - beg] = d;
    }
    eh[end].h = h1; eh[end].e = B2_MINUS_INF;
  }
  int score = eh[qlen].h;
  if (z && n_cigar_) {
    int n = 0, which = 0, k;
    i = tlen - 1;
    k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
      if (n + 2 > cigar_cap) { *n_cigar_ = -1; return score; }
      which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
      if (which == 0) { n = b2_push_cigar(cigar, n, 0, 1); --i; --k; }
      else if (which == 1) { n = b2_push_cigar(cigar, n, 2, 1); --i; }
      else { n = b2_push_cigar(cigar, n, 1, 1); --k; }
    }
    if (n + 2 > cigar_cap) { *n_cigar_ = -1; return score; }
    if (i >= 0) n = b2_push_cigar(cigar, n, 2, i + 1);
    if (k >= 0) n = b2_push_cigar(cigar, n, 1, k + 1);
    for (i = 0; i < n >> 1; ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
    *n_cigar_ = n;
  }
  return score;
}

// Striped local SW.  size: 1 = saturating bytes, 16 lanes; 2 = 16-bit words, 8 lanes.
// work: 4*slen*lanes int16 (H0,H1,E,Hmax); bbuf: tlen entries of (score<<32|row) scratch for the
// sub-optimal list (bounded by tlen because consecutive rows are merged).
B2_NOINLINE B2_HD inline B2Kswr b2_sw_local(int size, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int xtra,
                                int16_t* work, uint64_t* bbuf, unsigned long long* cells = 0) {
  const int lanes = size == 1 ? 16 : 8;
  const int slen = (qlen + lanes - 1) / lanes;
  const int n = slen * lanes;
  int shift = 127, mx = 0;
  for (int a = 0; a < 25; ++a) {
    if (mat[a] < (int8_t)shift) shift = mat[a];
    if (mat[a] > (int8_t)mx) mx = mat[a];
  }
  shift = (256 - shift) & 0xff;  // uint8 arithmetic: -min(mat)
  const int vmax = size == 1 ? 255 : 32767;
  const int oe_d = o_del + e_del, oe_i = o_ins + e_ins;
  int16_t *H0 = work, *H1 = work + n, *E = work + 2 * n, *Hmax = work + 3 * n;
  B2Kswr r; r.score = 0; r.te = r.qe = r.score2 = r.te2 = r.tb = r.qb = -1;
  int minsc = (xtra & B2_KSW_XSUBO) ? xtra & 0xffff : 0x10000;
  int endsc = (xtra & B2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
  int n_b = 0, te = -1, gmax = 0;
  for (int i = 0; i < n; ++i) H0[i] = E[i] = Hmax[i] = 0;
  // memory layout mirrors the SSE vectors: element (j, l) at j*lanes + l holds query column l*slen + j
  int fv[16], hv[16], mxv[16];
  for (int i = 0; i < tlen; ++i) {
    const int8_t* srow = mat + target[i] * 5;
    int imax;
    if (cells) *cells += (unsigned long long)qlen;
    for (int l = 0; l < lanes; ++l) { fv[l] = 0; mxv[l] = 0; }
    for (int l = lanes - 1; l > 0; --l) hv[l] = H0[(slen - 1) * lanes + l - 1];
    hv[0] = 0;
    for (int j = 0; j < slen; ++j) {
      for (int l = 0; l < lanes; ++l) {
        int col = l * slen + j;
        int s = col < qlen ? srow[query[col]] : 0;
        int h = hv[l], e = E[j * lanes + l], f = fv[l], t;
        if (size == 1) {
          h = h + (s + shift); if (h > 255) h = 255;
          h = h - shift; if (h < 0) h = 0;
        } else {
          h = h + s; if (h > 32767) h = 32767; if (h < -32768) h = -32768;
        }
        h = h > e ? h : e;
        h = h > f ? h : f;
        mxv[l] = mxv[l] > h ? mxv[l] : h;
        H1[j * lanes + l] = (int16_t)h;
        e = e - e_del; if (e < 0) e = 0;
        t = h - oe_d; if (t < 0) t = 0;
        e = e > t ? e : t;
        E[j * lanes + l] = (int16_t)e;
        f = f - e_ins; if (f < 0) f = 0;
        t = h - oe_i; if (t < 0) t = 0;
        fv[l] = f > t ? f : t;
        hv[l] = H0[j * lanes + l];
      }
    }
    if (o_ins >= 1 && e_ins >= 1 && !(xtra & B2_KSW_XLITERAL)) {
      // What the reference's lazy-F loop converges to (checked against the literal loop
      // below by tests/test_cpu.py::test_striped_sw_scan_equals_lazy_f): F entering a lane's column block is the
      // max-plus carry  x_l = max(own_l, x_{l-1} - slen*e_ins)  of the blocks to its left, and inside the block
      // it only decays (any H it raises re-opens a gap below what is already flowing).
      int x[16];
      for (int l = 0; l < lanes; ++l) x[l] = fv[l];
      for (int l = 1; l < lanes; ++l) { int c = x[l - 1] - slen * e_ins; if (c < 0) c = 0; x[l] = x[l] > c ? x[l] : c; }
      for (int l = lanes - 1; l > 0; --l) {
        int f = x[l - 1];
        for (int j = 0; j < slen && f > 0; ++j) {
          int hh = H1[j * lanes + l];
          H1[j * lanes + l] = (int16_t)(hh > f ? hh : f);
          f = f - e_ins; if (f < 0) f = 0;
        }
      }
    } else
    // lazy-F: propagate F across lane boundaries until it can no longer raise any H
    for (int k = 0, done = 0; k < 16 && !done; ++k) {
      for (int l = lanes - 1; l > 0; --l) fv[l] = fv[l - 1];
      fv[0] = 0;
      for (int j = 0; j < slen; ++j) {
        int all = 1;
        for (int l = 0; l < lanes; ++l) {
          int h = H1[j * lanes + l];
          h = h > fv[l] ? h : fv[l];
          H1[j * lanes + l] = (int16_t)h;
          h = h - oe_i; if (h < 0) h = 0;
          int f = fv[l] - e_ins; if (f < 0) f = 0;
          fv[l] = f;
          if (f > h) all = 0;
        }
        if (all) { done = 1; break; }
      }
    }
    imax = 0;
    for (int l = 0; l < lanes; ++l) imax = imax > mxv[l] ? imax : mxv[l];
    if (imax >= minsc) {
      if (n_b == 0 || (int32_t)bbuf[n_b - 1] + 1 != i) bbuf[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
      else if ((int)(bbuf[n_b - 1] >> 32) < imax) bbuf[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
    }
    if (imax > gmax) {
      gmax = imax; te = i;
      for (int j = 0; j < n; ++j) Hmax[j] = H1[j];
      if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
    }
    { int16_t* t = H1; H1 = H0; H0 = t; }
  }
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
  (void)vmax;
  return r;
}

// query/target are modified in place and restored (as the reference does).
B2_NOINLINE B2_HD inline B2Kswr b2_align2(int qlen, uint8_t* query, int tlen, uint8_t* target, const int8_t* mat, int o_del,
                              int e_del, int o_ins, int e_ins, int xtra, int16_t* work, uint64_t* bbuf,
                              unsigned long long* cells = 0) {
  const int size = (xtra & B2_KSW_XBYTE) ? 1 : 2;
  B2Kswr r = b2_sw_local(size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, work, bbuf, cells);
  if ((xtra & B2_KSW_XSTART) == 0 || ((xtra & B2_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
  b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target);
  B2Kswr rr = b2_sw_local(size, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins,
                          B2_KSW_XSTOP | r.score | (xtra & B2_KSW_XLITERAL), work, bbuf, cells);
  b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target);
  if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
  return r;
}
