// Chain -> alignment regions: reference fetch, seed extension driver, de-duplication and patching.
//
// Restates bwa/bntseq.c:349-446 (bns_pos2rid, bns_intv2rid, bns_get_seq, bns_fetch_seq),
// bwa/bwamem.c:621-628 (cal_max_gap), :632-786 (mem_chain2aln), :406-435 (mem_patch_reg),
// :437-489 (mem_sort_dedup_patch) and bwa/bwa.c:111-197 (bwa_gen_cigar2: band choice, gap-free
// shortcut, NM/MD).  Floating point appears only as IEEE double/float +,-,*,/ without contraction.
#pragma once
#include "b2_types.h"
#include "b2_sort.cuh"
#include "b2_ksw.cuh"
#include "b2_coop.cuh"
#include "b2_galn.cuh"
#include "b2_xext.cuh"

// Per-task bump scratch (stack discipline via mark/reset).
struct B2Scratch {
  uint8_t* base;
  size_t cap, used;
  int overflow;
  // algorithmic DP work of this lane: cells = sum over rows of the band actually computed
  unsigned long long ext_cells = 0, glob_cells = 0, sw_cells = 0;
  unsigned ext_calls = 0, glob_calls = 0, sw_calls = 0;
  // lane group serving this task (device: B2_TASK_LANES lanes running the same control flow)
  int lane = 0, gsize = 1;
  unsigned gmask = 0;
  const B2GalnCache* galn = 0;  // global alignments computed ahead of the finalisation kernel (b2_galn.cuh)
  long long* prof = 0;          // B200BWA_DEBUG_TIMES: cycles per phase of the finalisation, summed over the tasks
  long long prof_t = 0;
  B2_HD inline void tick(int k) {   // k < 0: start the clock
#if defined(__CUDA_ARCH__)
    if (prof) {
      const long long t = clock64();
      if (k >= 0 && lane == 0) atomicAdd((unsigned long long*)&prof[k], (unsigned long long)(t - prof_t));
      prof_t = t;
    }
#endif
  }
  B2_HD inline void gsync() const {
#if defined(__CUDA_ARCH__)
    if (gsize > 1) __syncwarp(gmask);
#endif
  }
  // sum of v over the lanes of the group (every lane gets the total)
  B2_HD inline int gsum(int v) const {
#if defined(__CUDA_ARCH__)
    for (int d = gsize >> 1; d; d >>= 1) v += __shfl_xor_sync(gmask, v, d, gsize);
#endif
    return v;
  }
  // bit i set iff lane i of the group passed a true predicate
  // OR of v over the lanes of the group (every lane gets the result)
  B2_HD inline unsigned gor(unsigned v) const {
#if defined(__CUDA_ARCH__)
    for (int d = gsize >> 1; d; d >>= 1) v |= __shfl_xor_sync(gmask, v, d, gsize);
#endif
    return v;
  }
  B2_HD inline unsigned gballot(int pred) const {
#if defined(__CUDA_ARCH__)
    if (gsize > 1) return (__ballot_sync(gmask, pred) & gmask) >> (__ffs(gmask) - 1);
#endif
    return pred ? 1u : 0u;
  }
  B2_HD inline void* alloc(size_t bytes, size_t align = 8) {
    size_t o = (used + align - 1) & ~(align - 1);
    if (o + bytes > cap) { overflow = 1; o = 0; if (bytes > cap) return base; }
    used = o + bytes;
    return base + o;
  }
  // Small second arena in shared memory (the finalisation kernels give every task group a few hundred bytes): the
  // private copy of the read and the reference window of an alignment are walked base by base several times (score,
  // NM/MD, reversal), and a round trip to L2 per step is what those loops otherwise cost.  Falls back to the global
  // arena when absent or full; marks cover both arenas.
  uint8_t* fast = 0;
  unsigned fast_cap = 0, fast_used = 0;
  // a task that does not fit the lane's (group's) own region is run again in a slot of the outlier pool (b2_retry_big)
  int big = 0;
  uint8_t* small_base = 0;
  size_t small_cap = 0;
  B2_HD inline void* alloc_fast(size_t bytes, size_t align = 8) {
    const size_t o = ((size_t)fast_used + align - 1) & ~(align - 1);
    if (fast && o + bytes <= fast_cap) { fast_used = (unsigned)(o + bytes); return fast + o; }
    return alloc(bytes, align);
  }
  B2_HD inline size_t mark() const { return used | (size_t)fast_used << 48; }
  B2_HD inline void reset(size_t m) { used = m & (((size_t)1 << 48) - 1); fast_used = (unsigned)(m >> 48); }
};

// the five characters of base codes 0..4 packed into one word: a shift instead of a table in memory
#define B2_BASES_FWD ((uint64_t)'A' | (uint64_t)'C' << 8 | (uint64_t)'G' << 16 | (uint64_t)'T' << 24 | (uint64_t)'N' << 32)
#define B2_BASES_REV ((uint64_t)'T' | (uint64_t)'G' << 8 | (uint64_t)'C' << 16 | (uint64_t)'A' << 24 | (uint64_t)'N' << 32)

// dst[i] = get(i) for i = lane, lane + gsize, ... < n: four elements in flight per lane -- the loads of a trip are issued
// before its stores (byte stores may alias anything as far as the compiler knows, which would otherwise serialise the
// loop at one memory round trip per element)
template <class F>
B2_HD inline void b2_gfill(uint8_t* dst, int n, int lane, int gsize, F get) {
  int i = lane;
  for (; i + 3 * gsize < n; i += 4 * gsize) {
    const uint8_t v0 = get(i), v1 = get(i + gsize), v2 = get(i + 2 * gsize), v3 = get(i + 3 * gsize);
    dst[i] = v0; dst[i + gsize] = v1; dst[i + 2 * gsize] = v2; dst[i + 3 * gsize] = v3;
  }
  for (; i < n; i += gsize) dst[i] = get(i);
}

// in-place reversal split over the lanes of the group (each lane swaps its own pairs)
B2_HD inline void b2_revseq_g(const B2Scratch& sc, int l, uint8_t* s) {
  const int h = l >> 1, g = sc.gsize;
  int i = sc.lane;
  for (; i + g < h; i += 2 * g) {   // two swaps in flight per lane
    const uint8_t a0 = s[i], b0 = s[l - 1 - i], a1 = s[i + g], b1 = s[l - 1 - i - g];
    s[i] = b0; s[l - 1 - i] = a0; s[i + g] = b1; s[l - 1 - i - g] = a1;
  }
  for (; i < h; i += g) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; }
}

// Dispatch of the three DP routines.  On the device every call runs a lane-group form: register-resident for whole
// warps (b2_coopreg.cuh), memory-resident row-parallel otherwise (b2_coop.cuh) -- a "group" may be a single lane
// (the lane-per-read kernels), its member mask is then the lane's own bit.  The host checker build (tests/hostsim)
// substitutes the scalar forms of tests/hostsim/b2_scalar_dp.cuh, which are not part of the library.
#if defined(__CUDA_ARCH__)
__device__ inline B2G b2_group_of(const B2Scratch& sc) {
  B2G g;
  g.size = sc.gsize; g.lane = sc.lane;
  g.mask = sc.gsize > 1 ? sc.gmask : (1u << (threadIdx.x & 31));
  return g;
}
#endif

B2_HD inline int b2_extend2_any(B2Scratch& sc, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus,
                                int zdrop, int h0, int* qle, int* tle, int* gtle, int* gscore, int* max_off, B2EH* eh) {
  ++sc.ext_calls;
#if defined(__CUDA_ARCH__)
  const B2G g = b2_group_of(sc);
  int s;
  if (sc.gsize > 1 && b2_extend2_reg_try(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, qle,
                                         tle, gtle, gscore, max_off, &sc.ext_cells, &s))
    return s;
  return b2_extend2_coop(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, qle, tle,
                         gtle, gscore, max_off, eh, &sc.ext_cells);
#elif defined(B2_HOSTSIM)
  return b2_extend2(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, qle, tle, gtle,
                    gscore, max_off, eh, &sc.ext_cells);
#else
  return 0;  // (host pass of the CUDA compiler: these functions only ever run on the device)
#endif
}

B2_HD inline int b2_global2_any(B2Scratch& sc, int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins, int w, B2EH* eh, uint8_t* z,
                                int* n_cigar_, uint32_t* cigar, int cigar_cap) {
  ++sc.glob_calls;
#if defined(__CUDA_ARCH__)
  const B2G g = b2_group_of(sc);
  int s;
  if (sc.gsize > 1 && b2_global2_diag_try(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, z, n_cigar_, cigar,
                                          cigar_cap, &sc.glob_cells, &s))
    return s;
  return b2_global2_coop(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, eh, z, n_cigar_, cigar,
                         cigar_cap, &sc.glob_cells);
#elif defined(B2_HOSTSIM)
  return b2_global2(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, w, eh, z, n_cigar_, cigar, cigar_cap,
                    &sc.glob_cells);
#else
  return 0;
#endif
}

// work: 4 * npad + 32 int16 (npad = query length rounded up to the 16 / 8 SSE lanes of ksw_u8 / ksw_i16)
B2_HD inline B2Kswr b2_align2_any(B2Scratch& sc, int qlen, uint8_t* query, int tlen, uint8_t* target, const int8_t* mat,
                                  int o_del, int e_del, int o_ins, int e_ins, int xtra, int16_t* work, uint64_t* bbuf) {
  ++sc.sw_calls;
#if defined(__CUDA_ARCH__)
  const B2G g = b2_group_of(sc);
  const int size = (xtra & B2_KSW_XBYTE) ? 1 : 2;
  // groups smaller than the SSE lane count (the 4-lane groups of the finalisation kernel, single lanes): blocks dealt
  // over the lanes, exchange through memory
  if (sc.gsize < 16) return b2_align2_g(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, work, bbuf, &sc.sw_cells);
  const bool reg_ok = size == 1 && sc.gsize == 16 && qlen <= 16 * B2_SW_SLMAX && mat[24] == mat[4] && mat[9] == mat[4];
  B2Kswr r = reg_ok ? b2_sw_local_u8_reg(g, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, bbuf, &sc.sw_cells)
                    : b2_sw_local_coop(g, size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, work, bbuf, &sc.sw_cells);
  if ((xtra & B2_KSW_XSTART) == 0 || ((xtra & B2_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
  // all lanes reverse the same private-to-the-group buffers: do it once
  if (sc.lane == 0) { b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target); }
  g.sync();
  B2Kswr rr = reg_ok ? b2_sw_local_u8_reg(g, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins,
                                          B2_KSW_XSTOP | r.score, bbuf, &sc.sw_cells)
                     : b2_sw_local_coop(g, size, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins,
                                        B2_KSW_XSTOP | r.score, work, bbuf, &sc.sw_cells);
  if (sc.lane == 0) { b2_revseq(r.qe + 1, query); b2_revseq(r.te + 1, target); }
  g.sync();
  if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
  return r;
#elif defined(B2_HOSTSIM)
  return b2_align2(qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, work, bbuf, &sc.sw_cells);
#else
  return B2Kswr();
#endif
}

B2_HD inline int b2_get_pac(const uint8_t* pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

B2_HD inline int64_t b2_depos(int64_t l_pac, int64_t pos, int* is_rev) {
  return (*is_rev = (pos >= l_pac)) ? (l_pac << 1) - 1 - pos : pos;
}

B2_HD inline int b2_pos2rid(const B2Index& ix, int64_t pos_f) {
  if (pos_f >= ix.l_pac) return -1;
  int left = 0, mid = 0, right = ix.n_seqs;
  while (left < right) {
    mid = (left + right) >> 1;
    if (pos_f >= ix.anns[mid].offset) {
      if (mid == ix.n_seqs - 1) break;
      if (pos_f < ix.anns[mid + 1].offset) break;
      left = mid + 1;
    } else right = mid;
  }
  return mid;
}

B2_HD inline int b2_intv2rid(const B2Index& ix, int64_t rb, int64_t re) {
  int is_rev;
  if (rb < ix.l_pac && re > ix.l_pac) return -2;
  int rid_b = b2_pos2rid(ix, b2_depos(ix.l_pac, rb, &is_rev));
  int rid_e = rb < re ? b2_pos2rid(ix, b2_depos(ix.l_pac, re - 1, &is_rev)) : rid_b;
  return rid_b == rid_e ? rid_b : -1;
}

// writes end-beg codes (after clamping) to seq; returns the length, 0 when bridging the strands
B2_HD inline int64_t b2_get_seq(int64_t l_pac, const uint8_t* pac, int64_t beg, int64_t end, uint8_t* seq) {
  if (end < beg) { int64_t t = beg; beg = end; end = t; }
  if (end > l_pac << 1) end = l_pac << 1;
  if (beg < 0) beg = 0;
  if (beg >= l_pac || end <= l_pac) {
    int64_t l = 0;
    if (beg >= l_pac) {
      int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
      for (int64_t k = end_f; k > beg_f; --k) seq[l++] = 3 - b2_get_pac(pac, k);
    } else {
      for (int64_t k = beg; k < end; ++k) seq[l++] = b2_get_pac(pac, k);
    }
    return end - beg;
  }
  return 0;
}

// clamp [*beg,*end) to the contig/strand holding mid; returns rid
B2_HD inline int b2_fetch_bounds(const B2Index& ix, int64_t* beg, int64_t mid, int64_t* end) {
  int is_rev;
  if (*end < *beg) { int64_t t = *beg; *beg = *end; *end = t; }
  int rid = b2_pos2rid(ix, b2_depos(ix.l_pac, mid, &is_rev));
  int64_t far_beg = ix.anns[rid].offset, far_end = far_beg + ix.anns[rid].len;
  if (is_rev) {
    int64_t tmp = far_beg;
    far_beg = (ix.l_pac << 1) - far_end;
    far_end = (ix.l_pac << 1) - tmp;
  }
  *beg = *beg > far_beg ? *beg : far_beg;
  *end = *end < far_end ? *end : far_end;
  return rid;
}

B2_HD inline int b2_cal_max_gap(const B2Opt& opt, int qlen) {
  int l_del = (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
  int l_ins = (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
  int l = l_del > l_ins ? l_del : l_ins;
  l = l > 1 ? l : 1;
  return l < opt.w << 1 ? l : opt.w << 1;
}

struct B2U64Lt { B2_HD bool operator()(uint64_t a, uint64_t b) const { return a < b; } };

// lane-parallel variant of b2_get_seq for a task group (every lane ends up seeing the whole window)
B2_NOINLINE B2_HD inline int64_t b2_get_seq_g(B2Scratch& sc, int64_t l_pac, const uint8_t* pac, int64_t beg, int64_t end, uint8_t* seq) {
  if (end < beg) { int64_t t = beg; beg = end; end = t; }
  if (end > l_pac << 1) end = l_pac << 1;
  if (beg < 0) beg = 0;
  if (beg >= l_pac || end <= l_pac) {
    const int64_t n = end - beg;
    sc.gsync();
    if (beg >= l_pac) {
      const int64_t end_f = (l_pac << 1) - 1 - beg;
      b2_gfill(seq, (int)n, sc.lane, sc.gsize, [&](int l) { return (uint8_t)(3 - b2_get_pac(pac, end_f - l)); });
    } else {
      b2_gfill(seq, (int)n, sc.lane, sc.gsize, [&](int l) { return (uint8_t)b2_get_pac(pac, beg + l); });
    }
    sc.gsync();
    return n;
  }
  return 0;
}

// Reference window and work buffers of one chain, set up on first use.
struct B2ChainPrep {
  int ready;
  int64_t rmax[2];
  uint8_t *rseq, *qs, *rs;
  B2EH* eh;
};

// reference window [rmax0, rmax1) of a chain (bwamem.c:641-660): every seed with the largest gaps, clamped to one
// strand and one contig
B2_HD inline void b2_chain_window(const B2Opt& opt, const B2Index& ix, int l_query, const B2Chain& c, const B2Seed* seeds,
                                  int64_t* rmax) {
  const int64_t l_pac = ix.l_pac;
  rmax[0] = l_pac << 1; rmax[1] = 0;
  for (int i = 0; i < c.n; ++i) {
    const B2Seed& t = seeds[i];
    int64_t b = t.rbeg - (t.qbeg + b2_cal_max_gap(opt, t.qbeg));
    int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + b2_cal_max_gap(opt, l_query - t.qbeg - t.len));
    rmax[0] = rmax[0] < b ? rmax[0] : b;
    rmax[1] = rmax[1] > e ? rmax[1] : e;
  }
  rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
  rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
  if (rmax[0] < l_pac && l_pac < rmax[1]) {
    if (seeds[0].rbeg < l_pac) rmax[1] = l_pac; else rmax[0] = l_pac;
  }
  b2_fetch_bounds(ix, &rmax[0], seeds[0].rbeg, &rmax[1]);
}

B2_NOINLINE B2_HD inline int b2_chain_prep(const B2Opt& opt, const B2Index& ix, int l_query, const B2Chain& c, const B2Seed* seeds,
                               B2ChainPrep& P, B2Scratch& sc) {
  if (P.ready) return 1;
  const int64_t l_pac = ix.l_pac;
  int64_t* rmax = P.rmax;
  b2_chain_window(opt, ix, l_query, c, seeds, rmax);
  const int64_t rlen = rmax[1] - rmax[0];
  P.rseq = (uint8_t*)sc.alloc((size_t)rlen + 1, 1);
  P.qs = (uint8_t*)sc.alloc((size_t)l_query + 1, 1);
  P.rs = (uint8_t*)sc.alloc((size_t)rlen + 1, 1);
  P.eh = (B2EH*)sc.alloc(sizeof(B2EH) * (size_t)(l_query + 2), 8);
  if (sc.overflow) return 0;
  b2_get_seq_g(sc, l_pac, ix.pac, rmax[0], rmax[1], P.rseq);
  P.ready = 1;
  return 1;
}

// Left/right extension of one seed inside its chain's window -> the region it yields (bwamem.c:709-781).
// Depends only on the seed, the chain and the read -- not on the regions found so far.
// xr: results of the inter-task pass for this seed ([0] left, [1] right; b2_xext.cuh) or null.  The reference window
// and the work buffers (P) are set up only if some DP has to run here.  Returns 0 on scratch overflow.
B2_NOINLINE B2_HD inline int b2_extend_seed(const B2Opt& opt, const B2Index& ix, int l_query, const uint8_t* query, const B2Chain& c,
                                const B2Seed* seeds, const B2Seed& s, B2ChainPrep& P, B2Reg& a, B2Scratch& sc,
                                const B2XextRes* xr = 0) {
  int max_off[2], aw[2], i;
  { B2Reg z = B2Reg(); a = z; }
  a.w = aw[0] = aw[1] = opt.w;
  a.score = a.truesc = -1;
  a.rid = c.rid;
  if (s.qbeg) {  // left extension on reversed prefixes
    int qle = 0, tle = 0, gtle = 0, gscore = 0;
    bool staged = false;
    for (i = 0; i < 2; ++i) {
      int prev = a.score;
      aw[0] = opt.w << i;
      if (i == 0 && xr && xr[0].w == aw[0] && xr[0].h0 == s.len * opt.a) {
        a.score = xr[0].score; qle = xr[0].qle; tle = xr[0].tle; gtle = xr[0].gtle; gscore = xr[0].gscore; max_off[0] = xr[0].max_off;
      } else {
        if (!b2_chain_prep(opt, ix, l_query, c, seeds, P, sc)) return 0;
        const int64_t tmp = s.rbeg - P.rmax[0];
        if (!staged) {
          sc.gsync();
          for (int k = sc.lane; k < s.qbeg; k += sc.gsize) P.qs[k] = query[s.qbeg - 1 - k];
          for (int64_t j = sc.lane; j < tmp; j += sc.gsize) P.rs[j] = P.rseq[tmp - 1 - j];
          sc.gsync();
          staged = true;
        }
        a.score = b2_extend2_any(sc, s.qbeg, P.qs, (int)tmp, P.rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[0],
                                 opt.pen_clip5, opt.zdrop, s.len * opt.a, &qle, &tle, &gtle, &gscore, &max_off[0], P.eh);
      }
      if (a.score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
    }
    if (gscore <= 0 || gscore <= a.score - opt.pen_clip5) {
      a.qb = s.qbeg - qle; a.rb = s.rbeg - tle;
      a.truesc = a.score;
    } else {
      a.qb = 0; a.rb = s.rbeg - gtle;
      a.truesc = gscore;
    }
  } else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }

  if (s.qbeg + s.len != l_query) {  // right extension
    int qle = 0, tle = 0, qe, gtle = 0, gscore = 0, sc0 = a.score;
    qe = s.qbeg + s.len;
    for (i = 0; i < 2; ++i) {
      int prev = a.score;
      aw[1] = opt.w << i;
      if (i == 0 && xr && xr[1].w == aw[1] && xr[1].h0 == sc0) {
        a.score = xr[1].score; qle = xr[1].qle; tle = xr[1].tle; gtle = xr[1].gtle; gscore = xr[1].gscore; max_off[1] = xr[1].max_off;
      } else {
        if (!b2_chain_prep(opt, ix, l_query, c, seeds, P, sc)) return 0;
        const int re = (int)(s.rbeg + s.len - P.rmax[0]);
        a.score = b2_extend2_any(sc, l_query - qe, query + qe, (int)(P.rmax[1] - P.rmax[0] - re), P.rseq + re, opt.mat, opt.o_del,
                                 opt.e_del, opt.o_ins, opt.e_ins, aw[1], opt.pen_clip3, opt.zdrop, sc0, &qle, &tle, &gtle,
                                 &gscore, &max_off[1], P.eh);
      }
      if (a.score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
    }
    if (gscore <= 0 || gscore <= a.score - opt.pen_clip3) {
      a.qe = qe + qle; a.re = s.rbeg + s.len + tle;
      a.truesc += a.score - sc0;
    } else {
      a.qe = l_query; a.re = s.rbeg + s.len + gtle;
      a.truesc += gscore - sc0;
    }
  } else { a.qe = l_query; a.re = s.rbeg + s.len; }

  a.seedcov = 0;
  for (i = 0; i < c.n; ++i) {
    const B2Seed& t = seeds[i];
    if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
  }
  a.w = aw[0] > aw[1] ? aw[0] : aw[1];
  a.seedlen0 = s.len;
  a.frac_rep = c.frac_rep;
  return 1;
}

// Extend the seeds of one chain, longest first, skipping seeds already covered by a region in av[] (bwamem.c:632-786).
//   mode 0  direct: av is the read's region list
//   mode 1  produce: av is a chain-local list (starts empty); every region computed is also stored in cand[seed],
//           cand_ok[seed] = 1 (per-chain tasks running ahead of the serial per-read pass)
//   mode 2  consume: av is the read's region list; a seed that has to be extended takes cand[seed] when present
//           and is computed in place otherwise
// srt: c.n u64 scratch.
B2_HD inline void b2_chain2aln(const B2Opt& opt, const B2Index& ix, int l_query, const uint8_t* query,
                               const B2Chain& c, const B2Seed* seeds, B2Reg* av, int* n_av, int cap_av,
                               uint64_t* srt, B2Scratch& sc, int mode = 0, B2Reg* cand = 0, uint8_t* cand_ok = 0,
                               const B2XextRes* xres = 0) {
  if (c.n == 0) return;
  const size_t mk = sc.mark();
  B2ChainPrep P;
  P.ready = 0;
  if (mode == 1) for (int i = sc.lane; i < c.n; i += sc.gsize) cand_ok[i] = 0;

  for (int i = 0; i < c.n; ++i) srt[i] = (uint64_t)seeds[i].score << 32 | (uint32_t)i;
  b2_introsort(srt, (long)c.n, B2U64Lt());

  for (int k = c.n - 1; k >= 0; --k) {
    const int si = (int)(uint32_t)srt[k];
    const B2Seed& s = seeds[si];
    int i;
    for (i = 0; i < *n_av; ++i) {  // was this seed already covered by an earlier extension?
      const B2Reg& p = av[i];
      int64_t rd;
      int qd, w, max_gap;
      if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
      if (s.len - p.seedlen0 > .1 * l_query) continue;
      qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
      max_gap = b2_cal_max_gap(opt, qd < rd ? qd : (int)rd);
      w = max_gap < p.w ? max_gap : p.w;
      if (qd - rd < w && rd - qd < w) break;
      qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
      max_gap = b2_cal_max_gap(opt, qd < rd ? qd : (int)rd);
      w = max_gap < p.w ? max_gap : p.w;
      if (qd - rd < w && rd - qd < w) break;
    }
    if (i < *n_av) {
      for (i = k + 1; i < c.n; ++i) {  // does an overlapping, off-diagonal seed in this chain disagree?
        if (srt[i] == 0) continue;
        const B2Seed& t = seeds[(uint32_t)srt[i]];
        if (t.len < s.len * .95) continue;
        if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
        if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
      }
      if (i == c.n) { srt[k] = 0; continue; }
    }
    if (*n_av >= cap_av) { sc.overflow = 1; break; }
    B2Reg& a = av[(*n_av)++];
    if (mode == 2 && cand_ok[si]) { a = cand[si]; continue; }
    if (!b2_extend_seed(opt, ix, l_query, query, c, seeds, s, P, a, sc, xres ? xres + 2 * si : 0)) break;
    if (mode == 1) {
      sc.gsync();
      if (sc.lane == 0) { cand[si] = a; cand_ok[si] = 1; }
    }
  }
  sc.gsync();
  sc.reset(mk);
}

// ---------------------------------------------------------------------------------------------
// Output of the end-point-constrained global alignment.
struct B2Cigar {
  uint32_t* cigar;  // capacity cap
  int cap;
  char* md;         // capacity md_cap
  int md_cap;
  int n_cigar, NM, l_md;
};

B2_HD inline int b2_put_int(char* s, int l, int cap, long c) {
  if (c >= 0 && c < 1000000000l) {   // MD run lengths: digits written in place
    const unsigned v = (unsigned)c;
    const int n = v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 :
                  v < 10000000u ? 7 : v < 100000000u ? 8 : 9;
    unsigned y = v;
    for (int k = n - 1; k >= 0; --k, y /= 10u) if (l + k < cap) s[l + k] = (char)('0' + y % 10u);
    return l + n;
  }
  char buf[24];
  int n = 0;
  unsigned long x = c < 0 ? (unsigned long)(-c) : (unsigned long)c;
  for (; x > 0; x /= 10) buf[n++] = (char)('0' + x % 10);
  if (c < 0) buf[n++] = '-';
  for (int i = n - 1; i >= 0; --i, ++l) if (l < cap) s[l] = buf[i];
  return l;
}

// want_cigar == 0: score only.  Returns 0 when the reference would have returned a null cigar
// without touching *score (reject conditions), 1 otherwise.
// pre: result of the speculative inter-task pass for exactly this (query, rb, re, w_), or null.
B2_NOINLINE B2_HD inline int b2_gen_cigar2(const B2Opt& opt, const B2Index& ix, int w_, int l_query, uint8_t* query, int64_t rb,
                               int64_t re, int* score, B2Cigar* out, B2Scratch& sc, const B2GalnRes* pre = 0) {
  const int64_t l_pac = ix.l_pac;
  const int8_t* mat = opt.mat;
  if (out) { out->n_cigar = 0; out->NM = -1; out->l_md = 0; if (out->md_cap > 0) out->md[0] = 0; }
  if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
  const size_t mk = sc.mark();
  uint8_t* rseq = (uint8_t*)sc.alloc_fast((size_t)(re - rb) + 1, 1);
  if (sc.overflow) { sc.reset(mk); return 0; }
  int64_t rlen = b2_get_seq_g(sc, l_pac, ix.pac, rb, re, rseq);
  if (re - rb != rlen) { sc.reset(mk); return 0; }
  if (rb >= l_pac) {  // reverse both so that indels are placed leftmost on the forward strand
    sc.gsync();
    b2_revseq_g(sc, l_query, query); b2_revseq_g(sc, (int)rlen, rseq);
    sc.gsync();
  }
  if (l_query == re - rb && w_ == 0) {
    if (out) { out->cigar[0] = (uint32_t)l_query << 4 | 0; out->n_cigar = 1; }
    int s = 0;
    for (int i = sc.lane; i < l_query; i += sc.gsize) s += mat[rseq[i] * 5 + query[i]];
    *score = sc.gsum(s);
  } else {
    const int w = b2_cigar_band(opt, l_query, (int)rlen, w_);
    if (pre && out && pre->n_cigar >= 0 && pre->n_cigar <= out->cap) {  // computed ahead by k_galn
      *score = pre->score;
      for (int k = sc.lane; k < pre->n_cigar; k += sc.gsize) out->cigar[k] = pre->cigar[k];
      out->n_cigar = pre->n_cigar;
      sc.gsync();
    } else {
    B2EH* eh = (B2EH*)sc.alloc(sizeof(B2EH) * (size_t)(l_query + 2), 8);
    uint8_t* z = 0;
    if (out) {
      int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
      z = (uint8_t*)sc.alloc((size_t)n_col * (size_t)rlen, 1);
    }
    if (sc.overflow) {
      if (rb >= l_pac) { sc.gsync(); b2_revseq_g(sc, l_query, query); sc.gsync(); }
      sc.reset(mk);
      return 0;
    }
    int nc = 0;
    sc.gsync();
    *score = b2_global2_any(sc, l_query, query, (int)rlen, rseq, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, eh, z,
                            out ? &nc : 0, out ? out->cigar : 0, out ? out->cap : 0);
    if (out) {
      if (nc < 0) { sc.overflow = 1; nc = 0; }
      out->n_cigar = nc;
    }
    }
  }
  if (out) {  // NM and MD
    int k, x, y, u, n_mm = 0, n_gap = 0, l = 0;
    const uint64_t int2base = rb < l_pac ? B2_BASES_FWD : B2_BASES_REV;
    char* md = out->md;
    // mismatches of a match run: a lane compares P consecutive positions, the group's results are ORed into one word
    // (small groups; whole warps vote instead), and the mismatches are then emitted in order
    const int P = sc.gsize <= 8 ? 4 : 1, chunk = P * sc.gsize;
    const int cap = out->md_cap - 1;
    for (k = 0, x = y = u = 0; k < out->n_cigar; ++k) {
      int op = out->cigar[k] & 0xf, len = (int)(out->cigar[k] >> 4);
      if (op == 0) {  // the lanes compare gsize positions at a time; mismatches are then emitted in order
        int cur = 0;
        for (int i0 = 0; i0 < len; i0 += chunk) {
          unsigned bits;
          if (P == 1) {
            const int i = i0 + sc.lane;
            bits = sc.gballot(i < len && query[x + i] != rseq[y + i]);
          } else {
            unsigned mine = 0;
            const int i1 = i0 + sc.lane * 4;
            for (int k = 0; k < 4; ++k) mine |= (unsigned)(i1 + k < len && query[x + i1 + k] != rseq[y + i1 + k]) << k;
            bits = sc.gor(mine << (sc.lane * 4));
          }
          while (bits) {
            int b = 0;
            while (!(bits >> b & 1)) ++b;
            bits &= bits - 1;
            const int p = i0 + b;
            u += p - cur;
            l = b2_put_int(md, l, cap, u);
            if (l < cap) md[l] = (char)(int2base >> (8 * rseq[y + p]));
            ++l; ++n_mm; u = 0;
            cur = p + 1;
          }
        }
        u += len - cur;
        x += len; y += len;
      } else if (op == 2) {
        if (k > 0 && k < out->n_cigar - 1) {
          l = b2_put_int(md, l, cap, u);
          if (l < cap) md[l] = '^';
          ++l;
          for (int i = 0; i < len; ++i) { if (l < cap) md[l] = (char)(int2base >> (8 * rseq[y + i])); ++l; }
          u = 0; n_gap += len;
        }
        y += len;
      } else if (op == 1) { x += len; n_gap += len; }
    }
    l = b2_put_int(md, l, cap, u);
    if (l > cap) { sc.overflow = 1; l = cap; }
    md[l] = 0;
    out->l_md = l;
    out->NM = n_mm + n_gap;
  }
  if (rb >= l_pac) { sc.gsync(); b2_revseq_g(sc, l_query, query); sc.gsync(); }
  sc.reset(mk);
  return 1;
}

// ---------------------------------------------------------------------------------------------
#define B2_PATCH_MAX_R_BW 0.05f
#define B2_PATCH_MIN_SC_RATIO 0.90f

// have_ref == 0 mirrors the calls that pass bns == pac == query == 0 (mate rescue), which never patch.
B2_HD inline int b2_patch_reg(const B2Opt& opt, const B2Index& ix, int have_ref, uint8_t* query, const B2Reg* a,
                              const B2Reg* b, int* _w, B2Scratch& sc) {
  int w, score = 0, q_s, r_s;
  double r;
  if (!have_ref || query == 0) return 0;
  if (a->rb < ix.l_pac && b->rb >= ix.l_pac) return 0;
  if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
  w = (int)((a->re - b->rb) - (a->qe - b->qb));
  w = w > 0 ? w : -w;
  r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
  r = r > 0. ? r : -r;
  if (a->re < b->rb || a->qe < b->qb) {
    if (w > opt.w << 1 || r >= B2_PATCH_MAX_R_BW) return 0;
  } else if (w > opt.w << 2 || r >= B2_PATCH_MAX_R_BW * 2) return 0;
  w += a->w + b->w;
  w = w < opt.w << 2 ? w : opt.w << 2;
  b2_gen_cigar2(opt, ix, w, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, 0, sc);
  q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
  r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
  if ((double)score / (q_s > r_s ? q_s : r_s) < B2_PATCH_MIN_SC_RATIO) return 0;
  *_w = w;
  return score;
}

struct B2RegReLt { B2_HD bool operator()(const B2Reg& a, const B2Reg& b) const { return a.re < b.re; } };
struct B2RegScLt {
  B2_HD bool operator()(const B2Reg& a, const B2Reg& b) const {
    return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb)));
  }
};

// Sorting the 96-byte region records of a repeat-rich read moves a lot of memory (the reads with dozens of regions are
// the tail of the finalisation kernel).  The permutation ks_introsort produces depends only on the outcomes of its
// comparisons, so the same algorithm run over 24-byte (key, index) records with the same predicate gives the same
// permutation; the records are then moved once.  One lane of the group sorts, all lanes copy.
struct B2SortKey { int64_t a; uint64_t b; int32_t c; int32_t idx; };
struct B2KeyRe { B2_HD B2SortKey operator()(const B2Reg& r) const { B2SortKey k; k.a = r.re; k.b = 0; k.c = 0; k.idx = 0; return k; } };
struct B2KeyReLt { B2_HD bool operator()(const B2SortKey& x, const B2SortKey& y) const { return x.a < y.a; } };
struct B2KeySc { B2_HD B2SortKey operator()(const B2Reg& r) const { B2SortKey k; k.a = r.rb; k.b = (uint64_t)(int64_t)r.qb; k.c = r.score; k.idx = 0; return k; } };
struct B2KeyScLt {
  B2_HD bool operator()(const B2SortKey& x, const B2SortKey& y) const {
    return x.c > y.c || (x.c == y.c && (x.a < y.a || (x.a == y.a && (int64_t)x.b < (int64_t)y.b)));
  }
};
#define B2_SORT_KEYS_MIN 5
template <class RegLt, class MakeKey, class KeyLt>
B2_HD inline void b2_sort_regs(B2Reg* a, int n, B2Scratch& sc, RegLt rlt, MakeKey mk, KeyLt klt) {
  const size_t need = (size_t)n * (sizeof(B2SortKey) + sizeof(B2Reg)) + 32;
  if (n < B2_SORT_KEYS_MIN || sc.used + need > sc.cap) { b2_introsort(a, (long)n, rlt); return; }
  const size_t mk0 = sc.mark();
  B2SortKey* keys = (B2SortKey*)sc.alloc(sizeof(B2SortKey) * (size_t)n, 8);
  B2Reg* tmp = (B2Reg*)sc.alloc(sizeof(B2Reg) * (size_t)n, 8);
  sc.gsync();
  for (int i = sc.lane; i < n; i += sc.gsize) { B2SortKey k = mk(a[i]); k.idx = i; keys[i] = k; tmp[i] = a[i]; }
  sc.gsync();
  if (sc.lane == 0) b2_introsort(keys, (long)n, klt);
  sc.gsync();
  for (int i = sc.lane; i < n; i += sc.gsize) a[i] = tmp[keys[i].idx];
  sc.gsync();
  sc.reset(mk0);
}

B2_NOINLINE B2_HD inline int b2_sort_dedup_patch(const B2Opt& opt, const B2Index& ix, int have_ref, uint8_t* query, int n,
                                     B2Reg* a, B2Scratch& sc) {
  int m, i, j;
  if (n <= 1) return n;
  b2_sort_regs(a, n, sc, B2RegReLt(), B2KeyRe(), B2KeyReLt());
  for (i = 0; i < n; ++i) a[i].n_comp = 1;
  for (i = 1; i < n; ++i) {
    B2Reg* p = &a[i];
    if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt.max_chain_gap) continue;
    for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt.max_chain_gap; --j) {
      B2Reg* q = &a[j];
      int64_t orr, oq, mr, mq;
      int score, w;
      if (q->qe == q->qb) continue;
      orr = q->re - p->rb;
      oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
      mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
      mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
      if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
        if (p->score < q->score) { p->qe = p->qb; break; }
        else q->qe = q->qb;
      } else if (q->rb < p->rb && (score = b2_patch_reg(opt, ix, have_ref, query, q, p, &w, sc)) > 0) {
        p->n_comp += q->n_comp + 1;
        p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
        p->sub = p->sub > q->sub ? p->sub : q->sub;
        p->csub = p->csub > q->csub ? p->csub : q->csub;
        p->qb = q->qb; p->rb = q->rb;
        p->truesc = p->score = score;
        p->w = w;
        q->qb = q->qe;
      }
    }
  }
  for (i = 0, m = 0; i < n; ++i)
    if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
  n = m;
  b2_sort_regs(a, n, sc, B2RegScLt(), B2KeySc(), B2KeyScLt());
  for (i = 1; i < n; ++i)
    if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
  for (i = 1, m = 1; i < n; ++i)
    if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
  return m;
}
