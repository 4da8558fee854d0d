// FM-index construction on the GPU: BWT + Occ + sampled SA of `bwa index`, byte for byte.
//
// What the reference computes (bwa/bwtindex.c:276-321): the BWT of T = forward + reverse-complement reference
// followed by '$' (IS for < 50 Mbp, BWT-SW above: two algorithms, one unique result), the cumulative A/C/G/T counts
// interleaved every 128 symbols (bwt_bwtupdate_core, bwa/bwtindex.c:151-173) and SA[k] for every 32nd row
// (bwt_cal_sa, bwa/bwt.c:62-84; 6.2 G serial LF steps at human scale).  Here the suffix array itself is built, group by
// group, and the BWT symbols and SA samples are read off it:
//
//   * T is packed 2 bits per base, 32 bases per 64-bit word; the sort key of suffix p at depth d is the word-aligned
//     window T[p+d, p+d+32) (zero = 'A' padding beyond the end; a suffix that runs out first sorts first, '$' < 'A').
//   * Suffixes are partitioned by their first 6 bases (4096 bins); consecutive bins form groups of at most
//     `group_max` suffixes, processed in lexicographic order, so memory is bounded by the group, not by the text:
//     6.2 G symbols need ~1.6 GB of text + 6.2 GB of output symbols + ~40 B per suffix of the current group.
//   * A group is sorted by a hand-written stable LSD radix sort (8-bit digits; one warp per tile ranks its elements
//     with match.any, no atomics on global memory); runs of equal keys are compacted and re-sorted by (run, next 32
//     bases) until every suffix is alone -- the planted repeats of the benchmark references need a few dozen rounds
//     over a shrinking subset.
//   * Row r of the sorted order gives BWT[r] = T[SA[r]-1] and, when r % 32 == 0, the SA sample -- no LF walk.
//
// The kernels use the launch shim of b2_backend.h, so the host checker build (tests/hostsim) runs the same bodies
// serially and the builder is compared with `bwa index` on the CPU-only test machine as well.
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "b2_backend.h"
#include "b2_types.h"
#include "b2_engine.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#if defined(B2_HOSTSIM)
#define IX_DEVICE_WARPS 0
#else
#define IX_DEVICE_WARPS 1
#endif

B2_HD inline u64 ix_key(const u64* W, int64_t p) {
  const int64_t w = p >> 5;
  const int s = (int)(p & 31) * 2;
  u64 k = W[w] << s;
  if (s) k |= W[w + 1] >> (64 - s);
  return k;
}
B2_HD inline int ix_valid(int64_t n, int64_t p) {
  const int64_t v = n - p;
  return v >= 32 ? 32 : (v < 0 ? 0 : (int)v);
}
B2_HD inline int ix_base(const u64* W, int64_t p) { return (int)(W[p >> 5] >> (62 - 2 * (int)(p & 31))) & 3; }

// T[i] for i in [0, 2L): forward strand, then its reverse complement (bwa/bntseq.c:296-302)
B2_KERNEL k_ix_text(const uint8_t* fwd, int64_t L, u64* W, int64_t n_words) {
  const int64_t n = 2 * L;
  for (int64_t w = B2_GTID(); w < n_words; w += B2_GSIZE()) {
    u64 v = 0;
    for (int b = 0; b < 32; ++b) {
      const int64_t i = (w << 5) + b;
      u64 c = 0;
      if (i < L) c = fwd[i] & 3;
      else if (i < n) c = 3 - (fwd[n - 1 - i] & 3);
      v = v << 2 | c;
    }
    W[w] = v;
  }
}

// number of suffixes per 6-base prefix
B2_KERNEL k_ix_hist6(const u64* W, int64_t n, u64* hist) {
#if IX_DEVICE_WARPS
  __shared__ u32 sh[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int64_t p = B2_GTID(); p < n; p += B2_GSIZE()) atomicAdd(&sh[ix_key(W, p) >> 52], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
#else
  for (int64_t p = B2_GTID(); p < n; p += B2_GSIZE()) hist[ix_key(W, p) >> 52] += 1;
#endif
}

// suffixes whose 6-base prefix lies in [lo, hi): (key at depth 0, position), in any order
B2_KERNEL k_ix_collect(const u64* W, int64_t n, int lo, int hi, u64* key, u64* pos, u64* count) {
  for (int64_t p0 = (int64_t)B2_GTID(); p0 < n + B2_WARP_LANES; p0 += B2_GSIZE()) {
    const int64_t p = p0;
    u64 k = 0;
    bool take = false;
    if (p < n) { k = ix_key(W, p); const int b = (int)(k >> 52); take = b >= lo && b < hi; }
#if IX_DEVICE_WARPS
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      u64 base = 0;
      if (lane == leader) base = atomicAdd(count, (u64)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (take) { const u64 j = base + __popc(m & ((1u << lane) - 1)); key[j] = k; pos[j] = (u64)p; }
    }
#else
    if (take) { const u64 j = (*count)++; key[j] = k; pos[j] = (u64)p; }
#endif
  }
}

// digit of element i for one radix pass: pass 0 = number of valid bases of the window (a suffix that ends inside the
// window sorts before its padded twins), passes 1..8 = bytes of the key, passes 9.. = bytes of the run id
B2_HD inline int ix_digit(int pass, u64 key, u64 pos, u32 seg, int64_t n, int64_t depth) {
  if (pass == 0) return ix_valid(n, (int64_t)pos + depth);
  if (pass <= 8) return (int)(key >> (8 * (pass - 1))) & 255;
  return (int)(seg >> (8 * (pass - 9))) & 255;
}

// per tile digit counts, hist[d * n_tiles + t]
B2_KERNEL k_ix_rs_hist(const u64* key, const u64* pos, const u32* seg, int64_t m, int pass, int64_t n, int64_t depth, u64* hist,
                       int64_t n_tiles, int tile) {
#if IX_DEVICE_WARPS
  extern __shared__ u32 shc[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u32* cnt = shc + wib * 256;
  const int64_t n_warps = (int64_t)B2_GSIZE() >> 5;
  for (int64_t t = (int64_t)B2_GTID() >> 5; t < n_tiles; t += n_warps) {
    for (int d = lane; d < 256; d += 32) cnt[d] = 0;
    __syncwarp();
    const int64_t b = t * tile, e = b + tile < m ? b + tile : m;
    for (int64_t i = b + lane; i < e; i += 32) atomicAdd(&cnt[ix_digit(pass, key[i], pos[i], seg ? seg[i] : 0, n, depth)], 1u);
    __syncwarp();
    for (int d = lane; d < 256; d += 32) hist[(int64_t)d * n_tiles + t] = cnt[d];
    __syncwarp();
  }
#else
  for (int64_t t = B2_GTID(); t < n_tiles; t += B2_GSIZE()) {
    u32 cnt[256];
    memset(cnt, 0, sizeof(cnt));
    const int64_t b = t * tile, e = b + tile < m ? b + tile : m;
    for (int64_t i = b; i < e; ++i) ++cnt[ix_digit(pass, key[i], pos[i], seg ? seg[i] : 0, n, depth)];
    for (int d = 0; d < 256; ++d) hist[(int64_t)d * n_tiles + t] = cnt[d];
  }
#endif
}

// stable scatter: offs = exclusive scan of hist
B2_KERNEL k_ix_rs_scatter(const u64* kin, const u64* pin, const u32* sin, u64* kout, u64* pout, u32* sout, int64_t m, int pass,
                          int64_t n, int64_t depth, const u64* offs, int64_t n_tiles, int tile) {
#if IX_DEVICE_WARPS
  extern __shared__ u64 sho[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u64* off = sho + wib * 256;
  const int64_t n_warps = (int64_t)B2_GSIZE() >> 5;
  for (int64_t t = (int64_t)B2_GTID() >> 5; t < n_tiles; t += n_warps) {
    for (int d = lane; d < 256; d += 32) off[d] = offs[(int64_t)d * n_tiles + t];
    __syncwarp();
    const int64_t b = t * tile, e = b + tile < m ? b + tile : m;
    for (int64_t i0 = b; i0 < e; i0 += 32) {
      const int64_t i = i0 + lane;
      const bool ok = i < e;
      u64 k = 0, p = 0;
      u32 s = 0;
      int d = -1 - lane;
      if (ok) { k = kin[i]; p = pin[i]; s = sin ? sin[i] : 0; d = ix_digit(pass, k, p, s, n, depth); }
      const unsigned mk = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(mk) - 1;
      u64 base = 0;
      if (lane == leader && ok) { base = off[d]; off[d] = base + __popc(mk); }
      base = __shfl_sync(0xffffffffu, base, leader);
      if (ok) {
        const u64 j = base + __popc(mk & ((1u << lane) - 1));
        kout[j] = k; pout[j] = p;
        if (sout) sout[j] = s;
      }
      __syncwarp();
    }
    __syncwarp();
  }
#else
  for (int64_t t = B2_GTID(); t < n_tiles; t += B2_GSIZE()) {
    u64 off[256];
    for (int d = 0; d < 256; ++d) off[d] = offs[(int64_t)d * n_tiles + t];
    const int64_t b = t * tile, e = b + tile < m ? b + tile : m;
    for (int64_t i = b; i < e; ++i) {
      const u32 s = sin ? sin[i] : 0;
      const u64 j = off[ix_digit(pass, kin[i], pin[i], s, n, depth)]++;
      kout[j] = kin[i]; pout[j] = pin[i];
      if (sout) sout[j] = s;
    }
  }
#endif
}

// in-place exclusive scan of a u64 array by one block (op 0: sum, op 1: running maximum, inclusive); total -> *tot
B2_KERNEL k_ix_scan(u64* a, int64_t n, int op, u64* tot) {
#if IX_DEVICE_WARPS
  __shared__ u64 part[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int64_t chunk = (n + T - 1) / T;
  const int64_t b = (int64_t)t * chunk, e = b + chunk < n ? b + chunk : n;
  u64 s = 0;
  for (int64_t i = b; i < e; ++i) s = op ? (s > a[i] ? s : a[i]) : s + a[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < T; d <<= 1) {
    u64 v = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] = op ? (part[t] > v ? part[t] : v) : part[t] + v;
    __syncthreads();
  }
  u64 run = t ? part[t - 1] : 0;
  for (int64_t i = b; i < e; ++i) {
    const u64 v = a[i];
    if (op) { run = run > v ? run : v; a[i] = run; }
    else { a[i] = run; run += v; }
  }
  if (t == T - 1 && tot) *tot = part[T - 1];
#else
  if (B2_GTID() != 0) return;
  u64 run = 0;
  for (int64_t i = 0; i < n; ++i) {
    const u64 v = a[i];
    if (op) { run = run > v ? run : v; a[i] = run; }
    else { a[i] = run; run += v; }
  }
  if (tot) *tot = run;
#endif
}

// Runs of equal (run id, key) among elements whose window is full are still tied.  tied[i] = 1 for their members,
// head[i] = i + 1 at the first member of a run (0 elsewhere) -- a running maximum of head then names each member's run.
B2_KERNEL k_ix_mark(const u64* key, const u64* pos, const u32* seg, int64_t m, int64_t n, int64_t depth, u64* tied, u64* head) {
  for (int64_t i = B2_GTID(); i < m; i += B2_GSIZE()) {
    const bool full = ix_valid(n, (int64_t)pos[i] + depth) == 32;
    bool tl = false, tr = false;
    if (full && i > 0) tl = key[i - 1] == key[i] && (!seg || seg[i - 1] == seg[i]) && ix_valid(n, (int64_t)pos[i - 1] + depth) == 32;
    if (full && i + 1 < m) tr = key[i + 1] == key[i] && (!seg || seg[i + 1] == seg[i]) && ix_valid(n, (int64_t)pos[i + 1] + depth) == 32;
    tied[i] = (tl || tr) ? 1 : 0;
    head[i] = (tr && !tl) ? (u64)i + 1 : 0;
  }
}

// Place the sorted elements (out[slot] = position) and compact the tied ones for the next round:
// their position, their final slot, and the slot of their run's first member as the new run id.
B2_KERNEL k_ix_place(const u64* pos, const u32* slot_in, int64_t m, const u64* tied, const u64* tidx, const u64* head, u64* out,
                     u64* tpos, u32* tslot, u32* tseg) {
  for (int64_t i = B2_GTID(); i < m; i += B2_GSIZE()) {
    const u64 sl = slot_in ? slot_in[i] : (u64)i;
    out[sl] = pos[i];
    if (tied[i]) {
      const u64 j = tidx[i];
      const int64_t h = (int64_t)head[i] - 1;  // index (this round's order) of the run's first member
      tpos[j] = pos[i];
      tslot[j] = (u32)sl;
      tseg[j] = slot_in ? slot_in[h] : (u32)h;
    }
  }
}

// rows row0 .. row0+m-1 of the suffix array: BWT symbol (4 marks the row of the whole text, '$') and SA samples
B2_KERNEL k_ix_emit(const u64* sa, int64_t m, int64_t row0, const u64* W, uint8_t* sym, u64* samp, int sa_shift, u64* primary) {
  for (int64_t i = B2_GTID(); i < m; i += B2_GSIZE()) {
    const int64_t r = row0 + i, p = (int64_t)sa[i];
    if (p == 0) { sym[r] = 4; *primary = (u64)r; }
    else sym[r] = (uint8_t)ix_base(W, p - 1);
    if ((r & ((1 << sa_shift) - 1)) == 0) samp[r >> sa_shift] = (u64)p;
  }
}

// 128 stored symbols (the '$' row removed) -> 8 data words of block b at payload[16 b + 8 ..] + its A/C/G/T counts
B2_KERNEL k_ix_pack(const uint8_t* sym, int64_t n, int64_t primary, u32* payload, u64* cnt, int64_t n_blocks) {
  for (int64_t b = B2_GTID(); b < n_blocks; b += B2_GSIZE()) {
    u64 c[4] = {0, 0, 0, 0};
    for (int w = 0; w < 8; ++w) {
      u32 v = 0;
      for (int k = 0; k < 16; ++k) {
        const int64_t idx = b * 128 + w * 16 + k;
        int s = 0;
        if (idx < n) { s = sym[idx + (idx >= primary ? 1 : 0)] & 3; ++c[s]; }
        v = v << 2 | (u32)s;
      }
      payload[b * 16 + 8 + w] = v;
    }
    for (int k = 0; k < 4; ++k) cnt[(int64_t)k * (n_blocks + 1) + b] = c[k];
  }
}

// cumulative counts before every block (after the exclusive scans of cnt)
B2_KERNEL k_ix_counts(const u64* cnt, u32* payload, int64_t n_blocks) {
  for (int64_t b = B2_GTID(); b < n_blocks; b += B2_GSIZE())
    for (int k = 0; k < 4; ++k) {
      const u64 v = cnt[(int64_t)k * (n_blocks + 1) + b];
      payload[b * 16 + 2 * k] = (u32)v; payload[b * 16 + 2 * k + 1] = (u32)(v >> 32);
    }
}

// key of every element at the current depth
B2_KERNEL k_ix_keys(const u64* W, const u64* pos, int64_t cnt, int64_t depth, u64* key) {
  for (int64_t i = B2_GTID(); i < cnt; i += B2_GSIZE()) key[i] = ix_key(W, (int64_t)pos[i] + depth);
}

// ---------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    if (b2_dev_malloc(&p, bytes + 64)) { p = nullptr; return 1; }
    cap = bytes;
    return 0;
  }
  void release() { b2_dev_free(p); p = nullptr; cap = 0; }
  ~DevBuf() { b2_dev_free(p); }
};
int ix_fail(std::string* err, const char* m) { if (err) *err = m; return 1; }
bool write_all(FILE* fp, const void* p, size_t n) { return fwrite(p, 1, n, fp) == n; }
}  // namespace

// fwd: l_pac base codes 0..3 (N already replaced).  Writes <prefix>.bwt and <prefix>.sa.
// group_max: suffixes sorted at a time (0: 2^29, ~45 GB of device memory at that size).
int b2_build_fm(const uint8_t* fwd, int64_t l_pac, const char* prefix, int device, int64_t group_max, int sa_intv, int verbose,
                std::string* err) {
  if (l_pac <= 0) return ix_fail(err, "empty reference");
  int sa_shift = 0;
  while ((1 << sa_shift) < sa_intv) ++sa_shift;
  if ((1 << sa_shift) != sa_intv) return ix_fail(err, "SA sample interval is not a power of 2");
#if !defined(B2_HOSTSIM)
  if (cudaSetDevice(device) != cudaSuccess) return ix_fail(err, "cudaSetDevice failed (no usable CUDA device)");
  if (cudaFuncSetAttribute(k_ix_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
    return ix_fail(err, "cudaFuncSetAttribute failed");
#endif
  (void)device;
  const int64_t n = 2 * l_pac;
  if (group_max <= 0) group_max = (int64_t)1 << 29;
  const int64_t n_words = (n >> 5) + 3;
  const b2_stream_t st = 0;
  const int GRID = 148 * 8, BLK = 256;
  DevBuf d_fwd, d_W, d_hist, d_cnt1, d_sym, d_samp, d_prim;
  if (d_fwd.ensure((size_t)l_pac) || d_W.ensure((size_t)n_words * 8) || d_hist.ensure(4096 * 8) || d_cnt1.ensure(16) ||
      d_sym.ensure((size_t)n + 2) || d_samp.ensure((size_t)((n >> sa_shift) + 2) * 8) || d_prim.ensure(8))
    return ix_fail(err, "device allocation failed (index construction)");
  b2_h2d(d_fwd.p, fwd, (size_t)l_pac, st);
  b2_dev_memset(d_hist.p, 0, 4096 * 8, st);
  b2_dev_memset(d_prim.p, 0xff, 8, st);
  B2_LAUNCH(k_ix_text, GRID, BLK, st, (const uint8_t*)d_fwd.p, l_pac, (u64*)d_W.p, n_words);
  B2_LAUNCH(k_ix_hist6, GRID, BLK, st, (const u64*)d_W.p, n, (u64*)d_hist.p);
  std::vector<u64> hist(4096);
  b2_d2h(hist.data(), d_hist.p, 4096 * 8, st);
  if (b2_sync(st)) return ix_fail(err, "device synchronisation failed (index construction)");
  d_fwd.release();
  const u64* W = (const u64*)d_W.p;
  uint8_t* sym = (uint8_t*)d_sym.p;
  {  // row 0 is the suffix '$' itself: SA = n, BWT = T[n-1] = complement of the first forward base
    const uint8_t s0 = (uint8_t)(3 - (fwd[0] & 3));
    b2_h2d(sym, &s0, 1, st);
    b2_sync(st);
  }
  u64 hmax = 0;
  for (u64 h : hist) hmax = std::max(hmax, h);
  if ((int64_t)hmax > group_max) group_max = (int64_t)hmax;  // one 6-base bin is the smallest unit
  if (group_max >= ((int64_t)1 << 32)) return ix_fail(err, "a single 6-base bin holds 2^32 suffixes or more");

  DevBuf K[2], P[2], S[2], SL[2], tied, head, tidx, out, rhist;
  const int TILE = 8192;
  int64_t row0 = 1;
  for (int lo = 0; lo < 4096;) {
    int hi = lo;
    int64_t m = 0;
    while (hi < 4096 && (m == 0 || m + (int64_t)hist[hi] <= group_max)) { m += (int64_t)hist[hi]; ++hi; }
    if (m == 0) { lo = hi; continue; }
    if (verbose >= 3) fprintf(stderr, "[M::b200bwa_index] suffixes starting with 6-mers %d..%d: %ld, rows from %ld\n", lo, hi - 1, (long)m, (long)row0);
    const size_t mm = (size_t)m;
    if (K[0].ensure(mm * 8) || K[1].ensure(mm * 8) || P[0].ensure(mm * 8) || P[1].ensure(mm * 8) || out.ensure(mm * 8) ||
        tied.ensure(mm * 8) || head.ensure(mm * 8) || tidx.ensure(mm * 8))
      return ix_fail(err, "device allocation failed (suffix group)");
    b2_dev_memset(d_cnt1.p, 0, 8, st);
    B2_LAUNCH(k_ix_collect, GRID, BLK, st, W, n, lo, hi, (u64*)K[0].p, (u64*)P[0].p, (u64*)d_cnt1.p);
    u64 got = 0;
    b2_d2h(&got, d_cnt1.p, 8, st);
    if (b2_sync(st) || (int64_t)got != m) return ix_fail(err, "suffix group size mismatch");

    // Rounds.  `cur` elements are still to be ordered (round 0: the whole group).  kc/pc/sc: key, position, run id in the
    // current order; slot[i]: final slot of the element at index i of that order (round 0: i itself).
    int64_t cur = m, depth = 0;
    u64 *kc = (u64*)K[0].p, *kn = (u64*)K[1].p, *pc = (u64*)P[0].p, *pn = (u64*)P[1].p;
    u32 *sc = nullptr, *sn = nullptr, *slot = nullptr;
    int slot_buf = 0;
    for (int round = 0; cur > 0; ++round) {
      if (round > 100000) return ix_fail(err, "suffix sort did not converge");
      const int64_t n_tiles = (cur + TILE - 1) / TILE;
      if (rhist.ensure((size_t)n_tiles * 256 * 8)) return ix_fail(err, "device allocation failed (radix histogram)");
      int seg_passes = 0;
      if (sc) { int64_t v = m - 1; while (v > 0) { ++seg_passes; v >>= 8; } }
      const int n_pass = 9 + seg_passes;
      const int rs_grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_tiles + 7) / 8, 148 * 8));
      for (int pass = 0; pass < n_pass; ++pass) {
        B2_LAUNCH_SMEM(k_ix_rs_hist, rs_grid, 256, 8 * 256 * sizeof(u32), st, (const u64*)kc, (const u64*)pc, (const u32*)sc, cur, pass, n,
                       depth, (u64*)rhist.p, n_tiles, TILE);
        B2_LAUNCH(k_ix_scan, 1, 1024, st, (u64*)rhist.p, n_tiles * 256, 0, (u64*)nullptr);
        B2_LAUNCH_SMEM(k_ix_rs_scatter, rs_grid, 256, 8 * 256 * sizeof(u64), st, (const u64*)kc, (const u64*)pc, (const u32*)sc, kn, pn, sn,
                       cur, pass, n, depth, (const u64*)rhist.p, n_tiles, TILE);
        std::swap(kc, kn); std::swap(pc, pn); std::swap(sc, sn);
      }
      B2_LAUNCH(k_ix_mark, GRID, BLK, st, (const u64*)kc, (const u64*)pc, (const u32*)sc, cur, n, depth, (u64*)tied.p, (u64*)head.p);
#if defined(B2_HOSTSIM)
      memcpy(tidx.p, tied.p, (size_t)cur * 8);
#else
      cudaMemcpyAsync(tidx.p, tied.p, (size_t)cur * 8, cudaMemcpyDeviceToDevice, st);
#endif
      b2_dev_memset(d_cnt1.p, 0, 8, st);
      B2_LAUNCH(k_ix_scan, 1, 1024, st, (u64*)tidx.p, cur, 0, (u64*)d_cnt1.p);
      B2_LAUNCH(k_ix_scan, 1, 1024, st, (u64*)head.p, cur, 1, (u64*)nullptr);
      u64 n_tied = 0;
      b2_d2h(&n_tied, d_cnt1.p, 8, st);
      if (b2_sync(st)) return ix_fail(err, "device synchronisation failed (suffix sort)");
      u32 *slot_next = nullptr, *seg_next = nullptr;
      if (n_tied) {
        if (SL[0].ensure(mm * 4) || SL[1].ensure(mm * 4) || S[0].ensure(mm * 4) || S[1].ensure(mm * 4))
          return ix_fail(err, "device allocation failed (tie refinement)");
        slot_buf = slot ? slot_buf ^ 1 : 0;   // the final slots move to the buffer the current ones are not in
        slot_next = (u32*)SL[slot_buf].p;
        seg_next = (u32*)S[0].p;              // (the current run ids are not read any more)
      }
      B2_LAUNCH(k_ix_place, GRID, BLK, st, (const u64*)pc, (const u32*)slot, cur, (const u64*)tied.p, (const u64*)tidx.p,
                (const u64*)head.p, (u64*)out.p, pn, slot_next, seg_next);
      if (n_tied == 0) break;
      depth += 32;
      cur = (int64_t)n_tied;
      std::swap(pc, pn);
      slot = slot_next;
      sc = seg_next;
      sn = (u32*)S[1].p;
      B2_LAUNCH(k_ix_keys, GRID, BLK, st, W, (const u64*)pc, cur, depth, kc);
    }
    B2_LAUNCH(k_ix_emit, GRID, BLK, st, (const u64*)out.p, m, row0, W, sym, (u64*)d_samp.p, sa_shift, (u64*)d_prim.p);
    if (b2_sync(st)) return ix_fail(err, "device synchronisation failed (emit)");
    row0 += m;
    lo = hi;
  }
  if (row0 != n + 1) return ix_fail(err, "suffix groups do not cover the text");
  for (DevBuf* b : {&K[0], &K[1], &P[0], &P[1], &S[0], &S[1], &SL[0], &SL[1], &tied, &head, &tidx, &out, &rhist}) b->release();

  // ---- BWT payload: 64-byte blocks {4 x u64 cumulative counts, 8 x u32 = 128 symbols}, then the closing totals ----
  u64 primary = 0;
  b2_d2h(&primary, d_prim.p, 8, st);
  if (b2_sync(st) || primary == (u64)-1) return ix_fail(err, "row of the whole text not found");
  const int64_t n_blocks = (n + 127) / 128, n_w16 = (n + 15) >> 4, wpad = n_blocks * 8 - n_w16;
  const int64_t payload_words = n_blocks * 16 - wpad + 8;
  DevBuf d_pay, d_cnt;
  if (d_pay.ensure((size_t)(n_blocks * 16 + 8) * 4) || d_cnt.ensure((size_t)(n_blocks + 1) * 4 * 8))
    return ix_fail(err, "device allocation failed (BWT layout)");
  b2_dev_memset(d_cnt.p, 0, (size_t)(n_blocks + 1) * 4 * 8, st);
  B2_LAUNCH(k_ix_pack, GRID, BLK, st, (const uint8_t*)sym, n, (int64_t)primary, (u32*)d_pay.p, (u64*)d_cnt.p, n_blocks);
  for (int k = 0; k < 4; ++k)
    B2_LAUNCH(k_ix_scan, 1, 1024, st, (u64*)d_cnt.p + (int64_t)k * (n_blocks + 1), n_blocks + 1, 0, (u64*)nullptr);
  B2_LAUNCH(k_ix_counts, GRID, BLK, st, (const u64*)d_cnt.p, (u32*)d_pay.p, n_blocks);
  u64 tot[4];
  for (int k = 0; k < 4; ++k) b2_d2h(&tot[k], (u64*)d_cnt.p + (int64_t)k * (n_blocks + 1) + n_blocks, 8, st);
  std::vector<u32> pay((size_t)payload_words);
  b2_d2h(pay.data(), d_pay.p, (size_t)(n_blocks * 16 - wpad) * 4, st);
  const int64_t n_sa = (n + sa_intv) / sa_intv;
  std::vector<u64> samp((size_t)n_sa);
  b2_d2h(samp.data(), d_samp.p, (size_t)n_sa * 8, st);
  if (b2_sync(st)) return ix_fail(err, "device synchronisation failed (BWT layout)");
  for (int k = 0; k < 4; ++k) { pay[(size_t)(n_blocks * 16 - wpad) + 2 * k] = (u32)tot[k]; pay[(size_t)(n_blocks * 16 - wpad) + 2 * k + 1] = (u32)(tot[k] >> 32); }
  u64 L2[5] = {0, 0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) L2[k + 1] = L2[k] + tot[k];
  if ((int64_t)L2[4] != n) return ix_fail(err, "symbol counts do not add up to the text length");
  {  // bwt_dump_bwt (bwa/bwt.c:385-394)
    FILE* fp = fopen((std::string(prefix) + ".bwt").c_str(), "wb");
    if (!fp) return ix_fail(err, "cannot write the .bwt file");
    bool ok = write_all(fp, &primary, 8) && write_all(fp, L2 + 1, 32) && write_all(fp, pay.data(), pay.size() * 4);
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return ix_fail(err, "write error on the .bwt file");
  }
  {  // bwt_dump_sa (bwa/bwt.c:396-405): sa[0] is not stored
    FILE* fp = fopen((std::string(prefix) + ".sa").c_str(), "wb");
    if (!fp) return ix_fail(err, "cannot write the .sa file");
    const u64 hdr2[2] = {(u64)sa_intv, (u64)n};
    bool ok = write_all(fp, &primary, 8) && write_all(fp, L2 + 1, 32) && write_all(fp, hdr2, 16) &&
              write_all(fp, samp.data() + 1, (size_t)(n_sa - 1) * 8);
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return ix_fail(err, "write error on the .sa file");
  }
  return 0;
}
