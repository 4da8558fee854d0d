// Finalisation: primary marking, mapping quality, CIGAR/NM/MD, XA, mate rescue, pairing and SAM
// text, one read (SE) or one pair (PE) per task.
//
// Restates bwa/bwamem.c:493-558 (mem_mark_primary_se), :945-969 (mem_approx_mapq_se), :792-799
// (infer_bw), :1057-1127 (mem_reg2aln), :801-939 (get_rlen, mem_aln2sam), :972-1017 (mem_reg2sam);
// bwa/bwamem_extra.c:90-140 (mem_gen_alt); bwa/bwamem_pair.c:23-44 (mem_infer_dir, cal_sub),
// :111-180 (mem_matesw), :182-243 (mem_pair), :249-388 (mem_sam_pe); utils.h:98-109 (hash_64).
// log() is only ever taken of integers: the host uploads a table computed with the same libm
// (SURVEY.md H8).  log(2*erfc(|ns|/sqrt2)) depends on the integer insert size and the batch's
// insert-size statistics: the host uploads one table per orientation and batch.
#pragma once
#include "b2_types.h"
#include "b2_align.cuh"

struct B2Tables {
  const double* logtab;   // logtab[i] = log((double)i), i < n_log
  int n_log;
  const double* pairtab[4];  // pairtab[d][dist - pes[d].low] = .721 * log(2 * erfc(|ns| * M_SQRT1_2)) * a
  int pair_n[4];
};

struct B2Read {
  const uint8_t* seq;   // codes 0..4
  const char* qual;     // null if FASTA
  const char* name;
  const char* comment;  // null unless -C and present
  int l_seq, l_name, l_comment;
};

struct B2Fmt {
  const char* rg_id;  // device copy of bwa_rg_id (may be empty)
  int l_rg;
};

struct B2Out {
  char* p;
  long cap, len;
  int lane = 0, gsize = 1;  // task group writing this record: bulk fields are written lane-strided
  // n characters tab[codes[i]] for i = from, from+step, ... (step = +1 or -1)
  // (four characters in flight per lane: loads first, then the stores -- see b2_gfill)
  // tab: the five characters for codes 0..4 packed into one word (B2_BASES_FWD / B2_BASES_REV), no table in memory
  B2_HD inline void put_codes(const uint8_t* codes, int from, int step, int n, uint64_t tab) {
    const int m = (long)n < cap - len ? n : (int)(cap - len > 0 ? cap - len : 0);   // characters that fit
    b2_gfill((uint8_t*)p + len, m, lane, gsize, [&](int i) { return (uint8_t)(tab >> (8 * codes[from + i * step])); });
    len += n;
  }
  B2_HD inline void put_chars(const char* s, int from, int step, int n) {
    const int m = (long)n < cap - len ? n : (int)(cap - len > 0 ? cap - len : 0);
    b2_gfill((uint8_t*)p + len, m, lane, gsize, [&](int i) { return (uint8_t)s[from + i * step]; });
    len += n;
  }
  B2_HD inline void putc_(int c) { if (len < cap) p[len] = (char)c; ++len; }
  B2_HD inline void putsn(const char* s, int l) { for (int i = 0; i < l; ++i) putc_(s[i]); }
  B2_HD inline void puts_(const char* s) { for (; *s; ++s) putc_(*s); }
  B2_HD inline void putl(long c) {
    if (c < 0) { putc_('-'); c = -c; }   // (LONG_MIN does not occur: positions, lengths, scores)
    unsigned long x = (unsigned long)c;
    if (x < 1000000000ul) {   // the common case in 32-bit arithmetic, digits written in place (no staging array)
      const unsigned v = (unsigned)x;
      const int n = v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 :
                    v < 10000000u ? 7 : v < 100000000u ? 8 : 9;
      unsigned y = v;
      for (int k = n - 1; k >= 0; --k, y /= 10u) if (len + k < cap) p[len + k] = (char)('0' + y % 10u);
      len += n;
      return;
    }
    char buf[24];
    int n = 0;
    for (; x > 0; x /= 10) buf[n++] = (char)('0' + x % 10);
    for (int i = n - 1; i >= 0; --i) putc_(buf[i]);
  }
};

B2_HD inline uint64_t b2_hash64(uint64_t key) {
  key += ~(key << 32); key ^= (key >> 22);
  key += ~(key << 13); key ^= (key >> 8);
  key += (key << 3);   key ^= (key >> 15);
  key += ~(key << 27); key ^= (key >> 31);
  return key;
}

B2_HD inline double b2_log_int(const B2Tables& tb, long v, int* err) {
  if (v < 0 || v >= tb.n_log) { *err = 1; return 0.; }
  return tb.logtab[v];
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// mem_flt_chained_seeds (bwa/bwamem.c:598-615) and mem_seed_sw (:571-596): for long reads, every seed of every kept
// chain is re-scored by a local SW of the seed +-50 bp and dropped when that score is too low.
#define B2_SHORT_EXT 50
#define B2_SHORT_LEN 200

// returns 0 when the reference returns before doing anything (short reads); else 1 and the minimum HSP score
B2_HD inline int b2_flt_seeds_threshold(const B2Opt& opt, const B2Tables& tb, int l_query, int* min_hsp, int* err) {
  const double min_l = opt.min_chain_weight ? (double)(1.1f * (float)opt.min_chain_weight) : (double)5.5f * b2_log_int(tb, l_query, err);
  *min_hsp = (int)(opt.a * min_l + .499);
  return !(min_l > (double)(0.05f * (float)l_query));
}

// -1: no need to test this seed
B2_HD inline int b2_seed_sw(const B2Opt& opt, const B2Index& ix, int l_query, const uint8_t* query, const B2Seed& s, B2Scratch& sc) {
  const int64_t l_pac = ix.l_pac;
  if (s.len >= B2_SHORT_LEN) return -1;
  int qb = s.qbeg, qe = s.qbeg + s.len;
  int64_t rb = s.rbeg, re = s.rbeg + s.len;
  const int64_t mid = (rb + re) >> 1;
  qb -= B2_SHORT_EXT; qb = qb > 0 ? qb : 0;
  qe += B2_SHORT_EXT; qe = qe < l_query ? qe : l_query;
  rb -= B2_SHORT_EXT; rb = rb > 0 ? rb : 0;
  re += B2_SHORT_EXT; re = re < l_pac << 1 ? re : l_pac << 1;
  if (rb < l_pac && l_pac < re) {
    if (mid < l_pac) re = l_pac;
    else rb = l_pac;
  }
  if (qe - qb >= B2_SHORT_LEN || re - rb >= B2_SHORT_LEN) return -1;
  b2_fetch_bounds(ix, &rb, mid, &re);  // bns_fetch_seq: clamp to the contig holding mid
  const size_t mk = sc.mark();
  const int ql = qe - qb, tl = (int)(re - rb);
  const int npad = ((ql + 7) / 8) * 8;
  uint8_t* q = (uint8_t*)sc.alloc((size_t)ql + 1, 1);
  uint8_t* ref = (uint8_t*)sc.alloc((size_t)tl + 1, 1);
  int16_t* work = (int16_t*)sc.alloc(sizeof(int16_t) * (4 * (size_t)npad + 32), 8);
  uint64_t* bbuf = (uint64_t*)sc.alloc(sizeof(uint64_t) * (size_t)(tl + 1), 8);
  if (sc.overflow) { sc.reset(mk); return -1; }
  sc.gsync();
  for (int k = sc.lane; k < ql; k += sc.gsize) q[k] = query[qb + k];
  b2_get_seq_g(sc, l_pac, ix.pac, rb, re, ref);
  sc.gsync();
  B2Kswr x = b2_align2_any(sc, ql, q, tl, ref, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, B2_KSW_XSTART, work, bbuf);
  sc.reset(mk);
  return x.score;
}

struct B2RegHashLt {
  B2_HD bool operator()(const B2Reg& a, const B2Reg& b) const {
    return a.score > b.score || (a.score == b.score && (a.is_alt < b.is_alt || (a.is_alt == b.is_alt && a.hash < b.hash)));
  }
};
struct B2RegHash2Lt {
  B2_HD bool operator()(const B2Reg& a, const B2Reg& b) const {
    return a.is_alt < b.is_alt || (a.is_alt == b.is_alt && (a.score > b.score || (a.score == b.score && a.hash < b.hash)));
  }
};
struct B2KeyHash { B2_HD B2SortKey operator()(const B2Reg& r) const { B2SortKey k; k.a = r.is_alt; k.b = r.hash; k.c = r.score; k.idx = 0; return k; } };
struct B2KeyHashLt {
  B2_HD bool operator()(const B2SortKey& x, const B2SortKey& y) const {
    return x.c > y.c || (x.c == y.c && (x.a < y.a || (x.a == y.a && x.b < y.b)));
  }
};
struct B2KeyHash2Lt {
  B2_HD bool operator()(const B2SortKey& x, const B2SortKey& y) const {
    return x.a < y.a || (x.a == y.a && (x.c > y.c || (x.c == y.c && x.b < y.b)));
  }
};

// z: int scratch of n entries
B2_HD inline void b2_mark_primary_core(const B2Opt& opt, int n, B2Reg* a, int* z) {
  int tmp = opt.a + opt.b, nz = 0;
  tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
  tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
  z[nz++] = 0;
  for (int i = 1; i < n; ++i) {
    int k;
    for (k = 0; k < nz; ++k) {
      int j = z[k];
      int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
      int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
      if (e_min > b_max) {
        int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
        if (e_min - b_max >= min_l * opt.mask_level) {
          if (a[j].sub == 0) a[j].sub = a[i].score;
          if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
          break;
        }
      }
    }
    if (k == nz) z[nz++] = i;
    else a[i].secondary = z[k];
  }
}

B2_NOINLINE B2_HD inline int b2_mark_primary_se(const B2Opt& opt, int n, B2Reg* a, int64_t id, int* z, B2Scratch& sc) {
  int i, n_pri;
  if (n == 0) return 0;
  for (i = n_pri = 0; i < n; ++i) {
    a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1;
    a[i].hash = b2_hash64((uint64_t)(id + i));
    if (!a[i].is_alt) ++n_pri;
  }
  b2_sort_regs(a, n, sc, B2RegHashLt(), B2KeyHash(), B2KeyHashLt());
  b2_mark_primary_core(opt, n, a, z);
  for (i = 0; i < n; ++i) {
    B2Reg* p = &a[i];
    p->secondary_all = i;
    if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
  }
  if (n_pri >= 0 && n_pri < n) {
    if (n_pri > 0) b2_sort_regs(a, n, sc, B2RegHash2Lt(), B2KeyHash(), B2KeyHash2Lt());
    for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
    for (i = 0; i < n; ++i) {
      if (a[i].secondary >= 0) {
        a[i].secondary_all = z[a[i].secondary];
        if (a[i].is_alt) a[i].secondary = 0x7fffffff;
      } else a[i].secondary_all = -1;
    }
    if (n_pri > 0) {
      for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
      b2_mark_primary_core(opt, n_pri, a, z);
    }
  } else {
    for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
  }
  return n_pri;
}

B2_HD inline int b2_approx_mapq_se(const B2Opt& opt, const B2Tables& tb, const B2Reg* a, int* err) {
  int mapq, l, sub = a->sub ? a->sub : opt.min_seed_len * opt.a;
  double identity;
  sub = a->csub > sub ? a->csub : sub;
  if (sub >= a->score) return 0;
  l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
  identity = 1. - (double)(l * opt.a - a->score) / (opt.a + opt.b) / l;
  if (a->score == 0) {
    mapq = 0;
  } else if (opt.mapQ_coef_len > 0) {
    double tmp;
    tmp = l < opt.mapQ_coef_len ? 1. : opt.mapQ_coef_fac / b2_log_int(tb, l, err);
    tmp *= identity * identity;
    mapq = (int)(6.02 * (a->score - sub) / opt.a * tmp * tmp + .499);
  } else {
    mapq = (int)(30.0 * (1. - (double)sub / a->score) * b2_log_int(tb, a->seedcov, err) + .499);
    mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
  }
  if (a->sub_n > 0) mapq -= (int)(4.343 * b2_log_int(tb, a->sub_n + 1, err) + .499);
  if (mapq > 60) mapq = 60;
  if (mapq < 0) mapq = 0;
  mapq = (int)(mapq * (1. - a->frac_rep) + .499);
  return mapq;
}

B2_HD inline int b2_infer_bw(int l1, int l2, int score, int a, int q, int r) {
  int w;
  if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
  w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
  int d = l1 - l2; d = d < 0 ? -d : d;
  if (w < d) w = d;
  return w;
}

// band of the first attempt of mem_reg2aln's loop (bwa/bwamem.c:1072-1080)
B2_HD inline int b2_reg2aln_w0(const B2Opt& opt, const B2Reg* ar) {
  const int qb = ar->qb, qe = ar->qe;
  const int64_t rb = ar->rb, re = ar->re;
  int tmp = b2_infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt.a, opt.o_del, opt.e_del);
  int w2 = b2_infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt.a, opt.o_ins, opt.e_ins);
  w2 = w2 > tmp ? w2 : tmp;
  if (w2 > opt.w) w2 = w2 < ar->w ? w2 : ar->w;
  return w2 < opt.w << 2 ? w2 : opt.w << 2;
}

B2_HD inline int b2_get_rlen(int n_cigar, const uint32_t* cigar) {
  int l = 0;
  for (int k = 0; k < n_cigar; ++k) { int op = cigar[k] & 0xf; if (op == 0 || op == 2) l += cigar[k] >> 4; }
  return l;
}

// ar == null: unmapped record.  qwork: l_query bytes of scratch holding a private copy of the read.
// The cigar/MD of the result live in `sc` (not released here).
B2_NOINLINE B2_HD inline B2Aln b2_reg2aln(const B2Opt& opt, const B2Index& ix, const B2Tables& tb, int l_query,
                              const uint8_t* query_, const B2Reg* ar, B2Scratch& sc, int* err) {
  B2Aln a;
  a.pos = 0; a.rid = 0; a.flag = 0; a.is_rev = a.is_alt = a.mapq = a.NM = 0; a.n_cigar = 0;
  a.cigar = 0; a.md = 0; a.XA = 0; a.score = a.sub = a.alt_sc = 0;
  if (ar == 0 || ar->rb < 0 || ar->re < 0) { a.rid = -1; a.pos = -1; a.flag |= 0x4; return a; }
  int i, w2, tmp, qb = ar->qb, qe = ar->qe, score = 0, is_rev, last_sc = -(1 << 30);
  int64_t pos, rb = ar->rb, re = ar->re;
  a.mapq = (uint32_t)(ar->secondary < 0 ? b2_approx_mapq_se(opt, tb, ar, err) : 0) & 0xff;
  if (ar->secondary >= 0) a.flag |= 0x100;
  w2 = b2_reg2aln_w0(opt, ar);  // (the loop's own clamp to 4*opt.w is idempotent)
  (void)tmp;
  B2Cigar cg;
  cg.cap = (qe - qb) + (int)(re - rb) + 4;
  cg.cigar = (uint32_t*)sc.alloc(sizeof(uint32_t) * (size_t)cg.cap, 4);
  cg.md_cap = 2 * (int)(re - rb) + (qe - qb) + 32;
  cg.md = (char*)sc.alloc((size_t)cg.md_cap, 1);
  uint8_t* query = (uint8_t*)sc.alloc_fast((size_t)l_query + 1, 1);
  if (sc.overflow) { a.rid = -1; a.pos = -1; a.flag |= 0x4; return a; }
  sc.gsync();
  b2_gfill(query, l_query, sc.lane, sc.gsize, [&](int k) { return query_[k]; });
  sc.gsync();
  cg.n_cigar = 0; cg.NM = -1; cg.l_md = 0; cg.md[0] = 0;
  i = 0;
  do {
    w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
    // (a gap-free region of band 0 runs no DP, bwa/bwa.c:131: nothing to look up)
    const B2GalnRes* pre = i == 0 && !(qe - qb == re - rb && w2 == 0) ? b2_galn_lookup(sc.galn, query_, qb, qe, rb, re, w2) : 0;
#if defined(B2_HOSTSIM) && defined(B2_GALN_STATS)
    { extern long b2_galn_stats[16]; const int wb = b2_cigar_band(opt, qe - qb, (int)(re - rb), w2);
      const bool nodp = (qe - qb == re - rb && w2 == 0);
      ++b2_galn_stats[nodp ? 0 : (pre ? 1 : (i > 0 ? 2 : (b2_galn_class(wb) < 0 ? 3 : 4)))];
      if (!nodp && !pre) b2_galn_stats[8] += (long)(2 * wb + 1 < qe - qb ? 2 * wb + 1 : qe - qb) * (re - rb);
      if (!nodp && !pre && i == 0 && getenv("B200BWA_DEBUG_GALN_MISS")) fprintf(stderr, "[D::miss] q %d-%d r %ld-%ld w2 %d wb %d truesc %d score %d arw %d sec %d\n", qb, qe, (long)rb, (long)re, w2, wb, ar->truesc, ar->score, ar->w, ar->secondary); }
#endif
    b2_gen_cigar2(opt, ix, w2, qe - qb, query + qb, rb, re, &score, &cg, sc, pre);
    if (score == last_sc || w2 == opt.w << 2) break;
    last_sc = score;
    w2 <<= 1;
  } while (++i < 3 && score < ar->truesc - opt.a);
  a.NM = (uint32_t)cg.NM & 0x3fffff;
  a.n_cigar = cg.n_cigar;
  a.cigar = cg.cigar;
  a.md = cg.md;
  pos = b2_depos(ix.l_pac, rb < ix.l_pac ? rb : re - 1, &is_rev);
  a.is_rev = (uint32_t)is_rev;
  if (a.n_cigar > 0) {  // squeeze out a leading or trailing deletion
    if ((a.cigar[0] & 0xf) == 2) {
      pos += a.cigar[0] >> 4;
      --a.n_cigar;
      for (i = 0; i < a.n_cigar; ++i) a.cigar[i] = a.cigar[i + 1];
    } else if ((a.cigar[a.n_cigar - 1] & 0xf) == 2) {
      --a.n_cigar;
    }
  }
  if (qb != 0 || qe != l_query) {  // clipping
    int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
    if (clip5) {
      for (i = a.n_cigar; i > 0; --i) a.cigar[i] = a.cigar[i - 1];
      a.cigar[0] = (uint32_t)clip5 << 4 | 3;
      ++a.n_cigar;
    }
    if (clip3) a.cigar[a.n_cigar++] = (uint32_t)clip3 << 4 | 3;
  }
  a.rid = b2_pos2rid(ix, pos);
  a.pos = pos - ix.anns[a.rid].offset;
  a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
  a.is_alt = (uint32_t)ar->is_alt & 1; a.alt_sc = ar->alt_sc;
  return a;
}

// ---------------------------------------------------------------------------------------------
// XA strings: XA[k] (k < n) -> NUL-terminated string in `sc` or null.
B2_HD inline int b2_pri_idx(double XA_drop_ratio, const B2Reg* a, int i) {
  int k = a[i].secondary_all;
  if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
  return -1;
}

B2_NOINLINE B2_HD inline char** b2_gen_alt(const B2Opt& opt, const B2Index& ix, const B2Tables& tb, int n, const B2Reg* a,
                               int l_query, const uint8_t* query, B2Scratch& sc, int* err) {
  if (n == 0) return 0;
  int* cnt = (int*)sc.alloc(sizeof(int) * (size_t)n, 4);
  char* has_alt = (char*)sc.alloc((size_t)n, 1);
  if (sc.overflow) return 0;
  int i, r, tot = 0;
  for (i = 0; i < n; ++i) { cnt[i] = 0; has_alt[i] = 0; }
  for (i = 0; i < n; ++i) {
    r = b2_pri_idx(opt.XA_drop_ratio, a, i);
    if (r >= 0) { ++cnt[r]; ++tot; if (a[i].is_alt) has_alt[r] = 1; }
  }
  if (tot == 0) return 0;
  char** XA = (char**)sc.alloc(sizeof(char*) * (size_t)n, 8);
  int* xlen = (int*)sc.alloc(sizeof(int) * (size_t)n, 4);
  if (sc.overflow) return 0;
  for (i = 0; i < n; ++i) { XA[i] = 0; xlen[i] = 0; }
  // two passes per primary would need the lengths first; instead each listed hit is formatted once
  // into a temporary and appended to its primary's growing buffer, which is sized by cnt[r].
  const int per_hit = 96 + 3 * l_query + ix.max_name_len;  // bound on one "name,+pos,cigar,NM;" entry (checked below)
  for (i = 0; i < n; ++i)
    if (cnt[i] > 0 && !(cnt[i] > opt.max_XA_hits_alt || (!has_alt[i] && cnt[i] > opt.max_XA_hits))) {
      XA[i] = (char*)sc.alloc((size_t)cnt[i] * per_hit + 1, 1);
      if (sc.overflow) return 0;
      XA[i][0] = 0;
    }
  for (i = 0; i < n; ++i) {
    if ((r = b2_pri_idx(opt.XA_drop_ratio, a, i)) < 0) continue;
    if (cnt[r] > opt.max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt.max_XA_hits)) continue;
    const size_t mk = sc.mark();
    B2Aln t = b2_reg2aln(opt, ix, tb, l_query, query, &a[i], sc, err);
    if (sc.overflow || t.rid < 0) { sc.overflow = 1; sc.reset(mk); return 0; }   // (no alignment: nothing below may index with it)
    B2Out o; o.p = XA[r] + xlen[r]; o.cap = (long)cnt[r] * per_hit - xlen[r]; o.len = 0;
    const B2Ann& an = ix.anns[t.rid];
    o.putsn(ix.names + an.name_off, an.name_len);
    o.putc_(','); o.putc_("+-"[t.is_rev]); o.putl(t.pos + 1);
    o.putc_(',');
    for (int k = 0; k < t.n_cigar; ++k) { o.putl(t.cigar[k] >> 4); o.putc_("MIDSHN"[t.cigar[k] & 0xf]); }
    o.putc_(','); o.putl((int)t.NM);
    o.putc_(';');
    if (o.len > o.cap) { sc.overflow = 1; sc.reset(mk); return 0; }
    xlen[r] += (int)o.len;
    XA[r][xlen[r]] = 0;
    sc.reset(mk);
  }
  return XA;
}

// ---------------------------------------------------------------------------------------------
// "%.3f" of score/alt_sc, rounding the exact binary value half-to-even like printf
B2_HD inline void b2_put_f3(B2Out& o, double x) {
  if (x < 0) { o.putc_('-'); x = -x; }
  union { double d; uint64_t u; } cv; cv.d = x;
  int e = (int)(cv.u >> 52 & 0x7ff);
  uint64_t m = cv.u & ((1ull << 52) - 1);
  if (e) m |= 1ull << 52; else e = 1;
  e -= 1075;  // x = m * 2^e
  unsigned long long q;  // round(x * 1000)
  if (e >= 0) q = (m << e) * 1000ull;   // only small magnitudes occur (a ratio of two alignment scores)
  else if (-e >= 64 + 10) q = 0;
  else {
    // m*1000 < 2^63; shift right by s = -e with round-half-even
    unsigned long long v = m * 1000ull;
    int s = -e;
    if (s >= 64) q = 0;
    else {
      q = v >> s;
      unsigned long long rem = v & ((1ull << s) - 1), half = 1ull << (s - 1);
      if (rem > half || (rem == half && (q & 1))) ++q;
    }
  }
  o.putl((long)(q / 1000));
  o.putc_('.');
  int f = (int)(q % 1000);
  o.putc_('0' + f / 100); o.putc_('0' + f / 10 % 10); o.putc_('0' + f % 10);
}

B2_NOINLINE B2_HD inline void b2_aln2sam(const B2Opt& opt, const B2Index& ix, const B2Fmt& fmt, B2Out& str, const B2Read& s, int n,
                             const B2Aln* list, int which, const B2Aln* m_) {
  B2Aln ptmp = list[which], *p = &ptmp, mtmp, *m = 0;
  int i;
  if (m_) { mtmp = *m_; m = &mtmp; }
  p->flag |= m ? 0x1 : 0;
  p->flag |= p->rid < 0 ? 0x4 : 0;
  p->flag |= m && m->rid < 0 ? 0x8 : 0;
  if (p->rid < 0 && m && m->rid >= 0) { p->rid = m->rid; p->pos = m->pos; p->is_rev = m->is_rev; p->n_cigar = 0; }
  if (m && m->rid < 0 && p->rid >= 0) { m->rid = p->rid; m->pos = p->pos; m->is_rev = p->is_rev; m->n_cigar = 0; }
  p->flag |= p->is_rev ? 0x10 : 0;
  p->flag |= m && m->is_rev ? 0x20 : 0;

  str.putsn(s.name, s.l_name); str.putc_('\t');
  str.putl((p->flag & 0xffff) | (p->flag & 0x10000 ? 0x100 : 0)); str.putc_('\t');
  if (p->rid >= 0) {
    const B2Ann& an = ix.anns[p->rid];
    str.putsn(ix.names + an.name_off, an.name_len); str.putc_('\t');
    str.putl(p->pos + 1); str.putc_('\t');
    str.putl((int)p->mapq); str.putc_('\t');
    if (p->n_cigar) {
      for (i = 0; i < p->n_cigar; ++i) {
        int c = p->cigar[i] & 0xf;
        if (!(opt.flag & B2_F_SOFTCLIP) && !p->is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
        str.putl(p->cigar[i] >> 4); str.putc_("MIDSH"[c]);
      }
    } else str.putc_('*');
  } else str.putsn("*\t0\t0\t*", 7);
  str.putc_('\t');

  if (m && m->rid >= 0) {
    if (p->rid == m->rid) str.putc_('=');
    else { const B2Ann& an = ix.anns[m->rid]; str.putsn(ix.names + an.name_off, an.name_len); }
    str.putc_('\t');
    str.putl(m->pos + 1); str.putc_('\t');
    if (p->rid == m->rid) {
      int64_t p0 = p->pos + (p->is_rev ? b2_get_rlen(p->n_cigar, p->cigar) - 1 : 0);
      int64_t p1 = m->pos + (m->is_rev ? b2_get_rlen(m->n_cigar, m->cigar) - 1 : 0);
      if (m->n_cigar == 0 || p->n_cigar == 0) str.putc_('0');
      else str.putl(-(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
    } else str.putc_('0');
  } else str.putsn("*\t0\t0", 5);
  str.putc_('\t');

  if (p->flag & 0x100) {
    str.putsn("*\t*", 3);
  } else if (!p->is_rev) {
    int qb = 0, qe = s.l_seq;
    if (p->n_cigar && which && !(opt.flag & B2_F_SOFTCLIP) && !p->is_alt) {
      if ((p->cigar[0] & 0xf) == 4 || (p->cigar[0] & 0xf) == 3) qb += p->cigar[0] >> 4;
      if ((p->cigar[p->n_cigar - 1] & 0xf) == 4 || (p->cigar[p->n_cigar - 1] & 0xf) == 3) qe -= p->cigar[p->n_cigar - 1] >> 4;
    }
    str.put_codes(s.seq, qb, 1, qe - qb, B2_BASES_FWD);
    str.putc_('\t');
    if (s.qual) str.put_chars(s.qual, qb, 1, qe - qb);
    else str.putc_('*');
  } else {
    int qb = 0, qe = s.l_seq;
    if (p->n_cigar && which && !(opt.flag & B2_F_SOFTCLIP) && !p->is_alt) {
      if ((p->cigar[0] & 0xf) == 4 || (p->cigar[0] & 0xf) == 3) qe -= p->cigar[0] >> 4;
      if ((p->cigar[p->n_cigar - 1] & 0xf) == 4 || (p->cigar[p->n_cigar - 1] & 0xf) == 3) qb += p->cigar[p->n_cigar - 1] >> 4;
    }
    str.put_codes(s.seq, qe - 1, -1, qe - qb, B2_BASES_REV);
    str.putc_('\t');
    if (s.qual) str.put_chars(s.qual, qe - 1, -1, qe - qb);
    else str.putc_('*');
  }

  if (p->n_cigar) {
    str.putsn("\tNM:i:", 6); str.putl((int)p->NM);
    str.putsn("\tMD:Z:", 6); str.puts_(p->md);
  }
  if (p->score >= 0) { str.putsn("\tAS:i:", 6); str.putl(p->score); }
  if (p->sub >= 0) { str.putsn("\tXS:i:", 6); str.putl(p->sub); }
  if (fmt.l_rg) { str.putsn("\tRG:Z:", 6); str.putsn(fmt.rg_id, fmt.l_rg); }
  if (!(p->flag & 0x100)) {
    for (i = 0; i < n; ++i)
      if (i != which && !(list[i].flag & 0x100)) break;
    if (i < n) {
      str.putsn("\tSA:Z:", 6);
      for (i = 0; i < n; ++i) {
        const B2Aln* r = &list[i];
        if (i == which || (r->flag & 0x100) || r->rid < 0) continue;
        const B2Ann& an = ix.anns[r->rid];
        str.putsn(ix.names + an.name_off, an.name_len); str.putc_(',');
        str.putl(r->pos + 1); str.putc_(',');
        str.putc_("+-"[r->is_rev]); str.putc_(',');
        for (int k = 0; k < r->n_cigar; ++k) { str.putl(r->cigar[k] >> 4); str.putc_("MIDSH"[r->cigar[k] & 0xf]); }
        str.putc_(','); str.putl((int)r->mapq);
        str.putc_(','); str.putl((int)r->NM);
        str.putc_(';');
      }
    }
    if (p->alt_sc > 0) { str.putsn("\tpa:f:", 6); b2_put_f3(str, (double)p->score / p->alt_sc); }
  }
  if (p->XA) { str.putsn("\tXA:Z:", 6); str.puts_(p->XA); }
  if (s.comment) { str.putc_('\t'); str.putsn(s.comment, s.l_comment); }
  if ((opt.flag & B2_F_REF_HDR) && p->rid >= 0 && ix.anns[p->rid].anno_len > 0) {
    const B2Ann& an = ix.anns[p->rid];
    str.putsn("\tXR:Z:", 6);
    for (i = 0; i < an.anno_len; ++i) { char c = ix.names[an.anno_off + i]; str.putc_(c == '\t' ? ' ' : c); }
  }
  str.putc_('\n');
}

// SE record set of one read (also the unpaired branch of PE).
B2_NOINLINE B2_HD inline void b2_reg2sam(const B2Opt& opt, const B2Index& ix, const B2Tables& tb, const B2Fmt& fmt, B2Out& str,
                             const B2Read& s, int n, B2Reg* a, int extra_flag, const B2Aln* m, B2Scratch& sc, int* err) {
  const size_t mk = sc.mark();
  char** XA = 0;
  if (!(opt.flag & B2_F_ALL)) XA = b2_gen_alt(opt, ix, tb, n, a, s.l_seq, s.seq, sc, err);
  B2Aln* aa = (B2Aln*)sc.alloc(sizeof(B2Aln) * (size_t)(n > 0 ? n : 1), 8);
  if (sc.overflow) { sc.reset(mk); return; }
  int n_aa = 0, k, l;
  for (k = l = 0; k < n; ++k) {
    B2Reg* p = &a[k];
    if (p->score < opt.T) continue;
    if (p->secondary >= 0 && (p->is_alt || !(opt.flag & B2_F_ALL))) continue;
    if (p->secondary >= 0 && p->secondary < 0x7fffffff && p->score < a[p->secondary].score * opt.drop_ratio) continue;
    B2Aln* q = &aa[n_aa++];
    *q = b2_reg2aln(opt, ix, tb, s.l_seq, s.seq, p, sc, err);
    q->XA = XA ? XA[k] : 0;
    q->flag |= extra_flag;
    if (p->secondary >= 0) q->sub = -1;
    if (l && p->secondary < 0) q->flag |= (opt.flag & B2_F_NO_MULTI) ? 0x10000 : 0x800;
    if (l && !p->is_alt && q->mapq > aa[0].mapq) q->mapq = aa[0].mapq;
    ++l;
  }
  if (sc.overflow) { sc.reset(mk); return; }   // an alignment above is missing (the task runs again with more scratch): no text from it
  if (n_aa == 0) {
    B2Aln t = b2_reg2aln(opt, ix, tb, s.l_seq, s.seq, 0, sc, err);
    t.flag |= extra_flag;
    b2_aln2sam(opt, ix, fmt, str, s, 1, &t, 0, m);
  } else {
    for (k = 0; k < n_aa; ++k) b2_aln2sam(opt, ix, fmt, str, s, n_aa, aa, k, m);
  }
  sc.reset(mk);
}

// ---------------------------------------------------------------------------------------------
// Paired-end
B2_HD inline int b2_infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t* dist) {
  int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
  int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
  *dist = p2 > b1 ? p2 - b1 : b1 - p2;
  return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

B2_HD inline int b2_cal_sub(const B2Opt& opt, int n, const B2Reg* a) {
  int j;
  for (j = 1; j < n; ++j) {
    int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
    int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
    if (e_min > b_max) {
      int min_l = a[j].qe - a[j].qb < a[0].qe - a[0].qb ? a[j].qe - a[j].qb : a[0].qe - a[0].qb;
      if (e_min - b_max >= min_l * opt.mask_level) break;
    }
  }
  return j < n ? a[j].score : opt.min_seed_len * opt.a;
}

// Insert-size candidate of one pair for mem_pestat (bwamem_pair.c:52-64).  Returns dir (0..3) or -1.
B2_HD inline int b2_pestat_candidate(const B2Opt& opt, int64_t l_pac, int n0, const B2Reg* r0, int n1, const B2Reg* r1,
                                     int64_t* is) {
  if (n0 == 0 || n1 == 0) return -1;
  if (b2_cal_sub(opt, n0, r0) > 0.8 * r0[0].score) return -1;
  if (b2_cal_sub(opt, n1, r1) > 0.8 * r1[0].score) return -1;
  if (r0[0].rid != r1[0].rid) return -1;
  int dir = b2_infer_dir(l_pac, r0[0].rb, r1[0].rb, is);
  if (*is && *is <= opt.max_ins) return dir;
  return -1;
}

// ---- mate-rescue windows ---------------------------------------------------------------------
// The SW inputs of a rescue attempt (mate sequence, reference window) depend only on the anchor hit, the
// orientation and the batch's insert-size statistics -- not on what earlier attempts found.  They are
// therefore planned for every pair up front, computed as independent tasks by a separate kernel (one lane
// group per task), and b2_matesw below looks the result up instead of running the SW inside the serial
// per-pair control flow.  An attempt that was not planned (possible only when an earlier rescue changed the
// mate's hit list in a way that un-skips an orientation) is still computed in place, so the outcome is the
// reference's either way.
struct B2RescTask {
  int32_t item;      // pair index inside the batch
  int16_t end_i, r;  // anchor read (0/1), orientation (0..3)
  int32_t j;         // anchor rank in the filtered list b[i]
  int32_t l_ms;
  int64_t rb, re;    // reference window after clamping
};
struct B2RescCache { const B2RescTask* tasks; const B2Kswr* res; int n; };

// Window of orientation r around anchor `a` (bwamem_pair.c:127-145).  Returns 1 iff the reference would run the SW.
B2_HD inline int b2_matesw_window(const B2Opt& opt, const B2Index& ix, const B2Pes pes[4], const B2Reg* a, int l_ms, int r,
                                  int64_t* rb_, int64_t* re_, int* is_rev_) {
  const int64_t l_pac = ix.l_pac;
  const int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1);
  int64_t rb, re;
  if (!is_rev) {
    rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
    re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
  } else {
    rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
    re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
  }
  if (rb < 0) rb = 0;
  if (re > l_pac << 1) re = l_pac << 1;
  *is_rev_ = is_rev;
  *rb_ = rb; *re_ = re;
  if (rb >= re) return 0;  // (the reference tests an uninitialised rid here; re - rb < min_seed_len rejects it anyway)
  int rid = b2_fetch_bounds(ix, &rb, (rb + re) >> 1, &re);
  *rb_ = rb; *re_ = re;
  return a->rid == rid && re - rb >= opt.min_seed_len;
}

B2_HD inline void b2_matesw_skip(const B2Index& ix, const B2Pes pes[4], const B2Reg* a, const B2Reg* ma, int n_ma, int skip[4]) {
  for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
  for (int i = 0; i < n_ma; ++i) {
    int64_t dist;
    int r = b2_infer_dir(ix.l_pac, a->rb, ma[i].rb, &dist);
    if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
  }
}

// Rescue attempts of one pair given the hit lists as they are before any rescue.  out == null: count only.
B2_HD inline int b2_rescue_plan(const B2Opt& opt, const B2Index& ix, const B2Pes pes[4], const int n[2], const B2Reg* const a[2],
                                const int l_seq[2], int item, B2RescTask* out) {
  if (opt.flag & B2_F_NO_RESCUE) return 0;
  int cnt = 0;
  for (int i = 0; i < 2; ++i) {
    int nb = 0;
    for (int j0 = 0; j0 < n[i] && nb < opt.max_matesw; ++j0) {
      if (a[i][j0].score < a[i][0].score - opt.pen_unpaired) continue;
      const int j = nb++;
      int skip[4];
      b2_matesw_skip(ix, pes, &a[i][j0], a[!i], n[!i], skip);
      for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        int64_t rb, re;
        int is_rev;
        if (!b2_matesw_window(opt, ix, pes, &a[i][j0], l_seq[!i], r, &rb, &re, &is_rev)) continue;
        if (out) {
          B2RescTask t;
          t.item = item; t.end_i = (int16_t)i; t.r = (int16_t)r; t.j = j; t.l_ms = l_seq[!i]; t.rb = rb; t.re = re;
          out[cnt] = t;
        }
        ++cnt;
      }
    }
  }
  return cnt;
}

// The SW of one planned attempt.  ms: the mate's base codes.
B2_HD inline B2Kswr b2_rescue_sw(const B2Opt& opt, const B2Index& ix, const B2RescTask& t, const uint8_t* ms, B2Scratch& sc) {
  const size_t mk = sc.mark();
  const int l_ms = t.l_ms;
  const int is_rev = (t.r >> 1 != (t.r & 1));
  const int tl = (int)(t.re - t.rb);
  const int xtra = B2_KSW_XSUBO | B2_KSW_XSTART | (l_ms * opt.a < 250 ? B2_KSW_XBYTE : 0) | (opt.min_seed_len * opt.a);
  const int lanes = (xtra & B2_KSW_XBYTE) ? 16 : 8;
  const int npad = ((l_ms + lanes - 1) / lanes) * lanes;
  uint8_t* seq = (uint8_t*)sc.alloc((size_t)l_ms + 1, 1);
  uint8_t* ref = (uint8_t*)sc.alloc((size_t)tl + 1, 1);
  int16_t* work = (int16_t*)sc.alloc(sizeof(int16_t) * (4 * (size_t)npad + 32), 8);
  uint64_t* bbuf = (uint64_t*)sc.alloc(sizeof(uint64_t) * (size_t)(tl + 1), 8);
  B2Kswr aln; aln.score = 0; aln.te = aln.qe = aln.score2 = aln.te2 = aln.tb = aln.qb = -1;
  if (sc.overflow) { sc.reset(mk); return aln; }
  sc.gsync();
  if (is_rev) { for (int i = sc.lane; i < l_ms; i += sc.gsize) seq[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4; }
  else { for (int i = sc.lane; i < l_ms; i += sc.gsize) seq[i] = ms[i]; }
  b2_get_seq_g(sc, ix.l_pac, ix.pac, t.rb, t.re, ref);
  sc.gsync();
  aln = b2_align2_any(sc, l_ms, seq, tl, ref, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, xtra, work, bbuf);
  sc.reset(mk);
  return aln;
}

// Mate rescue.  ma[*n_ma] may grow up to cap_ma.
B2_NOINLINE B2_HD inline int b2_matesw(const B2Opt& opt, const B2Index& ix, const B2Pes pes[4], const B2Reg* a, int l_ms,
                           const uint8_t* ms, B2Reg* ma, int* n_ma, int cap_ma, B2Scratch& sc, int end_i, int anchor_j,
                           const B2RescCache* cache) {
  const int64_t l_pac = ix.l_pac;
  int i, r, skip[4], n = 0;
  b2_matesw_skip(ix, pes, a, ma, *n_ma, skip);
  if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
  for (r = 0; r < 4; ++r) {
    if (skip[r]) continue;
    const size_t mk = sc.mark();
    int is_rev;
    int64_t rb, re;
    if (b2_matesw_window(opt, ix, pes, a, l_ms, r, &rb, &re, &is_rev)) {
      B2Kswr aln;
      int hit = 0;
      if (cache)
        for (int t = 0; t < cache->n; ++t)
          if (cache->tasks[t].end_i == end_i && cache->tasks[t].j == anchor_j && cache->tasks[t].r == r) { aln = cache->res[t]; hit = 1; break; }
      if (!hit) {
        B2RescTask t;
        t.item = 0; t.end_i = (int16_t)end_i; t.r = (int16_t)r; t.j = anchor_j; t.l_ms = l_ms; t.rb = rb; t.re = re;
        aln = b2_rescue_sw(opt, ix, t, ms, sc);
        if (sc.overflow) { sc.reset(mk); return n; }
      }
      if (aln.score >= opt.min_seed_len && aln.qb >= 0) {
        B2Reg b = B2Reg();
        b.rid = a->rid;
        b.is_alt = a->is_alt;
        b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
        b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
        b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
        b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
        b.score = aln.score;
        b.csub = aln.score2;
        b.secondary = -1;
        b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
        if (*n_ma >= cap_ma) { sc.overflow = 1; sc.reset(mk); return n; }
        ++*n_ma;
        for (i = 0; i < *n_ma - 1; ++i)
          if (ma[i].score < b.score) break;
        int tmp = i;
        for (i = *n_ma - 1; i > tmp; --i) ma[i] = ma[i - 1];
        ma[i] = b;
      }
      ++n;
    }
    if (n) *n_ma = b2_sort_dedup_patch(opt, ix, 0, 0, *n_ma, ma, sc);
    sc.reset(mk);
  }
  return n;
}

struct B2Pair64 { uint64_t x, y; };
struct B2Pair64Lt {
  B2_HD bool operator()(const B2Pair64& a, const B2Pair64& b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); }
};

// Best proper pair.  `id` is deliberately an int and `id<<8` wraps in 32 bits like the reference.
B2_NOINLINE B2_HD inline int b2_pair(const B2Opt& opt, const B2Index& ix, const B2Tables& tb, const B2Pes pes[4], const B2Reg* a0,
                         const B2Reg* a1, int id, int* sub, int* n_sub, int z[2], const int n_pri[2], B2Scratch& sc,
                         int* err) {
  const int64_t l_pac = ix.l_pac;
  const size_t mk = sc.mark();
  const int nv = n_pri[0] + n_pri[1];
  B2Pair64* v = (B2Pair64*)sc.alloc(sizeof(B2Pair64) * (size_t)(nv > 0 ? nv : 1), 8);
  if (sc.overflow) { sc.reset(mk); *sub = 0; *n_sub = 0; return 0; }
  int r, i, k, y[4], n = 0;
  for (r = 0; r < 2; ++r) {
    const B2Reg* a = r ? a1 : a0;
    for (i = 0; i < n_pri[r]; ++i) {
      const B2Reg* e = &a[i];
      B2Pair64 key;
      key.x = (uint64_t)(e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb);
      key.x = (uint64_t)e->rid << 32 | (key.x - (uint64_t)ix.anns[e->rid].offset);
      key.y = (uint64_t)e->score << 32 | (uint64_t)(i << 2) | (uint64_t)((e->rb >= l_pac) << 1) | (uint64_t)r;
      v[n++] = key;
    }
  }
  b2_introsort(v, (long)n, B2Pair64Lt());
  // the reference materialises every candidate pair in u[], sorts it and reads the top two; the order is
  // total, so the maximum, the runner-up and the count near the runner-up are tracked in two sweeps instead.
  B2Pair64 best, second;
  best.x = best.y = second.x = second.y = 0;
  long n_u = 0;
  const uint32_t id8 = (uint32_t)id << 8;
  const uint64_t idx = (uint64_t)(int64_t)(int32_t)id8;  // sign-extended like `id<<8` promoted to uint64
  int tmp = opt.a + opt.b;
  tmp = tmp > opt.o_del + opt.e_del ? tmp : opt.o_del + opt.e_del;
  tmp = tmp > opt.o_ins + opt.e_ins ? tmp : opt.o_ins + opt.e_ins;
  int subv = 0, cnt = 0;
  for (int pass = 0; pass < 2; ++pass) {
    y[0] = y[1] = y[2] = y[3] = -1;
    for (i = 0; i < n; ++i) {
      for (r = 0; r < 2; ++r) {
        int dir = r << 1 | (int)(v[i].y >> 1 & 1), which;
        if (pes[dir].failed) continue;
        which = r << 1 | (int)((v[i].y & 1) ^ 1);
        if (y[which] < 0) continue;
        for (k = y[which]; k >= 0; --k) {
          int64_t dist;
          int q;
          if ((int)(v[k].y & 3) != which) continue;
          dist = (int64_t)v[i].x - (int64_t)v[k].x;
          if (dist > pes[dir].high) break;
          if (dist < pes[dir].low) continue;
          long ti = (long)(dist - pes[dir].low);
          double term;
          if (ti < 0 || ti >= tb.pair_n[dir]) { *err = 1; term = 0.; } else term = tb.pairtab[dir][ti];
          q = (int)((double)((v[i].y >> 32) + (v[k].y >> 32)) + term + .499);
          if (q < 0) q = 0;
          B2Pair64 p;
          p.y = (uint64_t)k << 32 | (uint64_t)i;
          p.x = (uint64_t)q << 32 | (b2_hash64(p.y ^ idx) & 0xffffffffU);
          if (pass == 0) {
            if (n_u == 0) best = p;
            else if (B2Pair64Lt()(best, p)) { second = best; best = p; }
            else if (n_u == 1 || B2Pair64Lt()(second, p)) second = p;
            ++n_u;
          } else {
            if (!(p.x == best.x && p.y == best.y) && subv - (int)(p.x >> 32) <= tmp) ++cnt;
          }
        }
      }
      y[v[i].y & 3] = i;
    }
    if (pass == 0) {
      if (n_u == 0) break;
      subv = n_u > 1 ? (int)(second.x >> 32) : 0;
    }
  }
  int ret;
  if (n_u) {
    i = (int)(best.y >> 32); k = (int)(best.y << 32 >> 32);
    z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
    z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
    ret = (int)(best.x >> 32);
    *sub = subv;
    *n_sub = n_u > 1 ? cnt : 0;
  } else { ret = 0; *sub = 0; *n_sub = 0; }
  sc.reset(mk);
  return ret;
}

B2_HD inline int b2_raw_mapq(int diff, int a) { return (int)(6.02 * diff / a + .499); }

// One read pair: rescue, primary marking, pairing, SAM text for both mates.
// a[i] has capacity cap[i] and n[i] entries; id = global pair index.
B2_HD inline void b2_sam_pe(const B2Opt& opt, const B2Index& ix, const B2Tables& tb, const B2Fmt& fmt,
                            const B2Pes pes[4], uint64_t id, const B2Read s[2], B2Reg* a[2], int n[2], const int cap[2],
                            B2Out& out, B2Scratch& sc, int* err, const B2RescCache* cache = 0) {
  int i, j, z[2] = {0, 0}, o, subo = 0, n_sub = 0, extra_flag = 1, n_pri[2];
  B2Aln h[2];
  const size_t mk0 = sc.mark();
  sc.tick(-1);
  if (!(opt.flag & B2_F_NO_RESCUE)) {
    B2Reg* b[2];
    int nb[2] = {0, 0};
    for (i = 0; i < 2; ++i) {
      b[i] = (B2Reg*)sc.alloc(sizeof(B2Reg) * (size_t)(n[i] > 0 ? n[i] : 1), 8);
      if (sc.overflow) { sc.reset(mk0); return; }
      for (j = 0; j < n[i]; ++j)
        if (a[i][j].score >= a[i][0].score - opt.pen_unpaired) b[i][nb[i]++] = a[i][j];
    }
    for (i = 0; i < 2; ++i)
      for (j = 0; j < nb[i] && j < opt.max_matesw; ++j)
        b2_matesw(opt, ix, pes, &b[i][j], s[!i].l_seq, s[!i].seq, a[!i], &n[!i], cap[!i], sc, i, j, cache);
    sc.reset(mk0);
  }
  sc.tick(0);
  int zn = (n[0] > n[1] ? n[0] : n[1]) + 1;
  int* zbuf = (int*)sc.alloc(sizeof(int) * (size_t)zn, 4);
  if (sc.overflow) { sc.reset(mk0); return; }
  n_pri[0] = b2_mark_primary_se(opt, n[0], a[0], (int64_t)(id << 1 | 0), zbuf, sc);
  n_pri[1] = b2_mark_primary_se(opt, n[1], a[1], (int64_t)(id << 1 | 1), zbuf, sc);
  sc.tick(1);
  int no_pairing = 0;
  if (opt.flag & B2_F_NOPAIRING) no_pairing = 1;
  if (!no_pairing && n_pri[0] && n_pri[1] &&
      (o = b2_pair(opt, ix, tb, pes, a[0], a[1], (int)id, &subo, &n_sub, z, n_pri, sc, err)) > 0) {
    int is_multi[2], q_pe, score_un, q_se[2];
    for (i = 0; i < 2; ++i) {
      for (j = 1; j < n_pri[i]; ++j)
        if (a[i][j].secondary < 0 && a[i][j].score >= opt.T) break;
      is_multi[i] = j < n_pri[i] ? 1 : 0;
    }
    if (is_multi[0] || is_multi[1]) no_pairing = 1;
    if (!no_pairing) {
      score_un = a[0][0].score + a[1][0].score - opt.pen_unpaired;
      subo = subo > score_un ? subo : score_un;
      q_pe = b2_raw_mapq(o - subo, opt.a);
      if (n_sub > 0) q_pe -= (int)(4.343 * b2_log_int(tb, n_sub + 1, err) + .499);
      if (q_pe < 0) q_pe = 0;
      if (q_pe > 60) q_pe = 60;
      q_pe = (int)(q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
      if (o > score_un) {
        B2Reg* c[2];
        c[0] = &a[0][z[0]]; c[1] = &a[1][z[1]];
        for (i = 0; i < 2; ++i) {
          if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
          q_se[i] = b2_approx_mapq_se(opt, tb, c[i], err);
        }
        q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
        q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
        extra_flag |= 2;
        q_se[0] = q_se[0] < b2_raw_mapq(c[0]->score - c[0]->csub, opt.a) ? q_se[0] : b2_raw_mapq(c[0]->score - c[0]->csub, opt.a);
        q_se[1] = q_se[1] < b2_raw_mapq(c[1]->score - c[1]->csub, opt.a) ? q_se[1] : b2_raw_mapq(c[1]->score - c[1]->csub, opt.a);
      } else {
        z[0] = z[1] = 0;
        q_se[0] = b2_approx_mapq_se(opt, tb, &a[0][0], err);
        q_se[1] = b2_approx_mapq_se(opt, tb, &a[1][0], err);
      }
      for (i = 0; i < 2; ++i) {
        int k = a[i][z[i]].secondary_all;
        if (k >= 0 && k < n_pri[i]) {
          for (j = 0; j < n[i]; ++j)
            if (a[i][j].secondary_all == k || j == k) a[i][j].secondary_all = z[i];
          a[i][z[i]].secondary_all = -1;
        }
      }
      char** XA[2] = {0, 0};
      sc.tick(2);
      if (!(opt.flag & B2_F_ALL))
        for (i = 0; i < 2; ++i) XA[i] = b2_gen_alt(opt, ix, tb, n[i], a[i], s[i].l_seq, s[i].seq, sc, err);
      sc.tick(3);
      B2Aln g[2], aa[2][2];
      int n_aa[2] = {0, 0};
      for (i = 0; i < 2; ++i) {
        h[i] = b2_reg2aln(opt, ix, tb, s[i].l_seq, s[i].seq, &a[i][z[i]], sc, err);
        h[i].mapq = (uint32_t)q_se[i] & 0xff;
        h[i].flag |= 0x40 << i | extra_flag;
        h[i].XA = XA[i] ? XA[i][z[i]] : 0;
        aa[i][n_aa[i]++] = h[i];
        if (n_pri[i] < n[i]) {
          B2Reg* p = &a[i][n_pri[i]];
          if (p->score < opt.T || p->secondary >= 0 || !p->is_alt) continue;
          g[i] = b2_reg2aln(opt, ix, tb, s[i].l_seq, s[i].seq, p, sc, err);
          g[i].flag |= 0x800 | 0x40 << i | extra_flag;
          g[i].XA = XA[i] ? XA[i][n_pri[i]] : 0;
          aa[i][n_aa[i]++] = g[i];
        }
      }
      sc.tick(4);
      if (sc.overflow) { sc.reset(mk0); return; }   // an alignment above is missing: the task runs again with more scratch
      for (i = 0; i < n_aa[0]; ++i) b2_aln2sam(opt, ix, fmt, out, s[0], n_aa[0], aa[0], i, &h[1]);
      for (i = 0; i < n_aa[1]; ++i) b2_aln2sam(opt, ix, fmt, out, s[1], n_aa[1], aa[1], i, &h[0]);
      sc.tick(5);
      sc.reset(mk0);
      return;
    }
  }
  // no_pairing
  sc.tick(2);
  for (i = 0; i < 2; ++i) {
    int which = -1;
    if (n[i]) {
      if (a[i][0].score >= opt.T) which = 0;
      else if (n_pri[i] < n[i] && a[i][n_pri[i]].score >= opt.T) which = n_pri[i];
    }
    if (which >= 0) h[i] = b2_reg2aln(opt, ix, tb, s[i].l_seq, s[i].seq, &a[i][which], sc, err);
    else h[i] = b2_reg2aln(opt, ix, tb, s[i].l_seq, s[i].seq, 0, sc, err);
  }
  if (sc.overflow) { sc.reset(mk0); return; }
  if (!(opt.flag & B2_F_NOPAIRING) && h[0].rid == h[1].rid && h[0].rid >= 0) {
    int64_t dist;
    int d = b2_infer_dir(ix.l_pac, a[0][0].rb, a[1][0].rb, &dist);
    if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
  }
  sc.tick(4);
  b2_reg2sam(opt, ix, tb, fmt, out, s[0], n[0], a[0], 0x41 | extra_flag, &h[1], sc, err);
  b2_reg2sam(opt, ix, tb, fmt, out, s[1], n[1], a[1], 0x81 | extra_flag, &h[0], sc, err);
  sc.tick(6);
  sc.reset(mk0);
}
