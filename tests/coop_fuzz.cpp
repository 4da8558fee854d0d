// Equivalence fuzz of the register-resident group-cooperative DP (b2_coopreg.cuh) against the scalar forms of
// b2_ksw.cuh (which the host checker build pins to the oracle), with the lane group emulated by coroutines
// (simt_emu.h).  Usage: coop_fuzz <trials> [seed].  Prints "... mismatches N"; exit status 1 on any mismatch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define B2_WITH_SCALAR_DP 1   // the scalar reference side: tests/hostsim/b2_scalar_dp.cuh
#include "b2_ksw.cuh"
#include "b2_coopreg.cuh"
#include "b2_galn.cuh"
#include "b2_xext.cuh"
#include "simt_emu.h"

static const int GS = 32;  // lanes per group, as the device runs them (whole warps)

static void mutate(const std::vector<uint8_t>& t, int st, int qlen, double sub, double ind, std::vector<uint8_t>& q) {
  q.resize(qlen);
  int p = st;
  const int tlen = (int)t.size();
  for (int i = 0; i < qlen; ++i) {
    double u = drand48();
    if (u < ind) { q[i] = lrand48() % 4; continue; }
    if (u < 2 * ind) p += 1 + lrand48() % 6;
    uint8_t c = p < tlen ? t[p] : lrand48() % 4; ++p;
    if (drand48() < sub) c = (c + 1 + lrand48() % 3) & 3;
    if (drand48() < 0.004) c = 4;
    q[i] = c;
  }
}

template <int C>
static int run_ext(EmuWarp& W, int qlen, const uint8_t* q, int tlen, const uint8_t* t, const int8_t* mat, int od, int ed, int oi,
                   int ei, int w, int eb, int zd, int h0, int out[5], unsigned long long* cells) {
  int res[GS], o[GS][5];
  unsigned long long cl[GS] = {0};
  W.run(GS, [&](int lane) {
    EmuG g; g.w = &W; g.lane = lane; g.size = GS; g.mask = 0xffffffffu;
    res[lane] = b2_extend2_reg<C>(g, qlen, q, tlen, t, mat, od, ed, oi, ei, w, eb, zd, h0, &o[lane][0], &o[lane][1], &o[lane][2],
                                  &o[lane][3], &o[lane][4], &cl[lane]);
  });
  for (int l = 1; l < GS; ++l)
    if (res[l] != res[0] || memcmp(o[l], o[0], sizeof(o[0]))) return -999999;  // lanes must agree
  memcpy(out, o[0], sizeof(o[0]));
  *cells = cl[0];
  return res[0];
}

template <int C>
static int run_glob(EmuWarp& W, int qlen, const uint8_t* q, int tlen, const uint8_t* t, const int8_t* mat, int od, int ed, int oi,
                    int ei, int w, uint8_t* z, int* nc, uint32_t* cigar, int cap, unsigned long long* cells) {
  int res[GS], ncl[GS];
  unsigned long long cl[GS] = {0};
  W.run(GS, [&](int lane) {
    EmuG g; g.w = &W; g.lane = lane; g.size = GS; g.mask = 0xffffffffu;
    res[lane] = b2_global2_diag<C>(g, qlen, q, tlen, t, mat, od, ed, oi, ei, w, z, nc ? &ncl[lane] : 0, cigar, cap, &cl[lane]);
  });
  for (int l = 1; l < GS; ++l)
    if (res[l] != res[0] || (nc && ncl[l] != ncl[0])) return -999999;
  if (nc) *nc = ncl[0];
  *cells = cl[0];
  return res[0];
}

int main(int argc, char** argv) {
  int trials = argc > 1 ? atoi(argv[1]) : 3000;
  srand48(argc > 2 ? atol(argv[2]) : 4242);
  EmuWarp W;
  long bad_e = 0, bad_g = 0, n_e = 0, n_g = 0, n_l = 0, bad_l = 0, n_x = 0, bad_x = 0;
  for (int tr = 0; tr < trials; ++tr) {
    int a = 1 + lrand48() % 3, b = 1 + lrand48() % 6;
    int o_del = lrand48() % 9, e_del = 1 + lrand48() % 3, o_ins = lrand48() % 9, e_ins = 1 + lrand48() % 3;
    if (tr % 3 == 0) { a = 1; b = 4; o_del = o_ins = 6; e_del = e_ins = 1; }
    int8_t mat[25]; int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? a : -b; mat[k++] = -1; }
    for (int j = 0; j < 5; ++j) mat[k++] = -1;
    if (!b2_mat_simple(mat)) { printf("matrix shape check failed\n"); return 1; }
    {  // ---- extension ----
      int qlen = lrand48() % 5 == 0 ? (int)(lrand48() % 20) : (int)(1 + lrand48() % 250);
      int tlen = qlen + (int)(lrand48() % (qlen / 2 + 30));
      if (lrand48() % 10 == 0) tlen = 1 + lrand48() % 40;
      std::vector<uint8_t> t(tlen), q;
      for (auto& c : t) c = lrand48() % 4;
      if (lrand48() % 8 == 0) { q.resize(qlen); for (auto& c : q) c = lrand48() % 4; }
      else mutate(t, 0, qlen, 0.01 * (lrand48() % 12), 0.004 * (lrand48() % 6), q);
      if (lrand48() % 20 == 0) for (auto& c : t) if (lrand48() % 30 == 0) c = 4;
      int wopts[] = {1, 3, 10, 30, 100, 200};
      int w = wopts[lrand48() % 6], zdrop = lrand48() % 4 == 0 ? 0 : 20 + lrand48() % 120, h0 = 1 + lrand48() % 150;
      int end_bonus = lrand48() % 3 == 0 ? 0 : 5;
      std::vector<B2EH> eh(qlen + 2);
      int r0[5], r1[5] = {0, 0, 0, 0, 0};
      unsigned long long c0 = 0, c1 = 0;
      int s0 = b2_extend2(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, &r0[0], &r0[1],
                          &r0[2], &r0[3], &r0[4], eh.data(), &c0);
      int s1;
      const int need = (qlen + 1 + GS - 1) / GS;  // the device's choice of cells per lane (b2_extend2_reg_try)
      if (need <= 1) s1 = run_ext<1>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r1, &c1);
      else if (need <= 2) s1 = run_ext<2>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r1, &c1);
      else if (need <= 3) s1 = run_ext<3>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r1, &c1);
      else if (need <= 5) s1 = run_ext<5>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r1, &c1);
      else s1 = run_ext<8>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, r1, &c1);
      ++n_e;
      if (h0 + qlen * a < 65000) {  // the inter-task (one extension per lane) form of b2_xext.cuh, lane-strided storage
        const int zs = 1 + (int)(lrand48() % 3);
        std::vector<uint32_t> ehv((size_t)(qlen + 2) * zs + 8), qv((size_t)(qlen / 8 + 2) * zs + 8), tv((size_t)(tlen / 8 + 2) * zs + 8);
        B2XMem M; M.eh = ehv.data(); M.q = qv.data(); M.t = tv.data(); M.stride = zs;
        b2_xext_stage(M.q, zs, qlen, [&](int k) { return (int)q[k]; });
        b2_xext_stage(M.t, zs, tlen, [&](int k) { return (int)t[k]; });
        int r2[5]; unsigned long long c2 = 0;
        int s2 = b2_xext_one(M, qlen, tlen, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, &r2[0], &r2[1], &r2[2], &r2[3], &r2[4], &c2);
        ++n_x;
        if (s2 != s0 || memcmp(r0, r2, sizeof(r0)) || c2 != c0) {
          if (++bad_x <= 5) fprintf(stderr, "XEXT mismatch trial %d qlen %d tlen %d w %d h0 %d: %d vs %d\n", tr, qlen, tlen, w, h0, s0, s2);
        }
        if (h0 + qlen * a < 255) {  // the 8 + 8 bit cell form (batches whose scores stay below 256)
          std::vector<uint16_t> eh8((size_t)(qlen + 2) * zs + 8);
          B2XMemT<uint16_t> M8; M8.eh = eh8.data(); M8.q = qv.data(); M8.t = tv.data(); M8.stride = zs;
          int r3[5]; unsigned long long c3 = 0;
          int s3 = b2_xext_one(M8, qlen, tlen, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, &r3[0], &r3[1], &r3[2], &r3[3], &r3[4], &c3);
          ++n_x;
          if (s3 != s0 || memcmp(r0, r3, sizeof(r0)) || c3 != c0) {
            if (++bad_x <= 5) fprintf(stderr, "XEXT8 mismatch trial %d qlen %d tlen %d w %d h0 %d: %d vs %d\n", tr, qlen, tlen, w, h0, s0, s3);
          }
        }
      }
      if (s0 != s1 || memcmp(r0, r1, sizeof(r0)) || c0 != c1) {
        if (++bad_e <= 5)
          fprintf(stderr, "EXT mismatch trial %d qlen %d tlen %d w %d h0 %d zdrop %d: %d (%d %d %d %d %d) cells %llu vs %d (%d %d %d %d %d) cells %llu\n", tr,
                  qlen, tlen, w, h0, zdrop, s0, r0[0], r0[1], r0[2], r0[3], r0[4], c0, s1, r1[0], r1[1], r1[2], r1[3], r1[4], c1);
      }
    }
    {  // ---- global ----
      if (o_ins == 0) o_ins = 1;
      if (o_del == 0) o_del = 1;
      int qlen = 1 + lrand48() % 250;
      int dl = lrand48() % 3 == 0 ? 0 : (int)(lrand48() % 12) - 6;
      int tlen = qlen + dl; if (tlen < 1) tlen = 1;
      std::vector<uint8_t> t(tlen), q;
      for (auto& c : t) c = lrand48() % 4;
      if (lrand48() % 10 == 0) { q.resize(qlen); for (auto& c : q) c = lrand48() % 4; }
      else mutate(t, 0, qlen, 0.01 * (lrand48() % 10), 0.004 * (lrand48() % 5), q);
      int adl = abs(qlen - tlen);
      int w = adl + (int)(lrand48() % 4 == 0 ? lrand48() % 60 : lrand48() % 8);
      if (lrand48() % 4) w = w > adl + 3 ? w : adl + 3;
      if (w > 63) w = 63;
      if (w < adl) continue;
      const bool want_cigar = lrand48() % 5 != 0;
      int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
      std::vector<uint8_t> z0((size_t)n_col * tlen + 1, 0xee), z1((size_t)n_col * tlen + 1, 0xee);
      int cap = qlen + tlen + 4;
      if (lrand48() % 50 == 0) cap = 3;
      std::vector<uint32_t> cg0(cap + 2, 0), cg1(cap + 2, 0);
      std::vector<B2EH> eh(qlen + 2);
      int nc0 = 0, nc1 = 0;
      unsigned long long c0 = 0, c1 = 0;
      int s0 = b2_global2(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, eh.data(), want_cigar ? z0.data() : 0,
                          want_cigar ? &nc0 : 0, cg0.data(), cap, &c0);
      int s1;
      const int need = (2 * w + 1 + GS - 1) / GS;  // (b2_global2_diag_try)
      uint8_t* zz = want_cigar ? z1.data() : 0;
      int* ncp = want_cigar ? &nc1 : 0;
      if (need <= 1) s1 = run_glob<1>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zz, ncp, cg1.data(), cap, &c1);
      else if (need <= 2) s1 = run_glob<2>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zz, ncp, cg1.data(), cap, &c1);
      else s1 = run_glob<4>(W, qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zz, ncp, cg1.data(), cap, &c1);
      ++n_g;
      bool bad = s0 != s1 || nc0 != nc1 || c0 != c1;
      if (!bad && want_cigar && nc0 > 0 && memcmp(cg0.data(), cg1.data(), sizeof(uint32_t) * nc0)) bad = true;
      if (!bad && want_cigar && z0 != z1) bad = true;
      // the inter-task (one alignment per lane) form of b2_galn.cuh, with a lane stride as on the device
      if (!bad && want_cigar && cap > 3 && b2_galn_class(w) >= 0) {
        B2GalnRes gr; gr.n_cigar = -7; gr.score = 0;
        const int zs = 1 + (int)(lrand48() % 3);
        std::vector<uint32_t> zw((size_t)(tlen + 1) * B2_GALN_MAX_ZWORDS * zs + 8, 0xdeadbeef);
        const int cls = b2_galn_class(w);
        if (cls == 0) b2_galn_one<9>(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zw.data(), zs, &gr);
        else if (cls == 1) b2_galn_one<17>(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zw.data(), zs, &gr);
        else if (cls == 2) b2_galn_one<33>(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zw.data(), zs, &gr);
        else b2_galn_one<65>(qlen, q.data(), tlen, t.data(), mat, o_del, e_del, o_ins, e_ins, w, zw.data(), zs, &gr);
        ++n_l;
        if (gr.score != s0) bad = true;
        if (gr.n_cigar >= 0) { if (gr.n_cigar != nc0 || memcmp(gr.cigar, cg0.data(), sizeof(uint32_t) * nc0)) bad = true; }
        else if (nc0 <= B2_GALN_MAXCIG) bad = true;  // may only give up when the CIGAR does not fit
        if (bad) ++bad_l;
      }
      if (bad && ++bad_g <= 5)
        fprintf(stderr, "GLOB mismatch trial %d qlen %d tlen %d w %d cigar %d: score %d nc %d cells %llu vs %d %d %llu\n", tr, qlen, tlen, w,
                (int)want_cigar, s0, nc0, c0, s1, nc1, c1);
    }
  }
  printf("lane-per-extension trials %ld mismatches %ld; lane-per-alignment trials %ld mismatches %ld; ", n_x, bad_x, n_l, bad_l);
  printf("extend trials %ld mismatches %ld; global trials %ld mismatches %ld\n", n_e, bad_e, n_g, bad_g);
  return (bad_e || bad_g || bad_l || bad_x) ? 1 : 0;
}
