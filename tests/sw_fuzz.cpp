// Equivalence fuzz: striped local SW with the max-plus carry scan (what the engine runs) vs the reference's
// lazy-F loop emulated literally (B2_KSW_XLITERAL), over random scorings, byte and word mode, related and
// unrelated sequences, ambiguous bases, XSUBO/XSTART and XSTOP modes.  Prints "mismatches N".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define B2_WITH_SCALAR_DP 1   // the scalar reference side: tests/hostsim/b2_scalar_dp.cuh
#include "b2_ksw.cuh"
#include "b2_coopreg.cuh"
#include "simt_emu.h"

// the generic-group form the device uses for groups smaller than the SSE lane count, with the group emulated
static B2Kswr run_group(EmuWarp& W, int G, int qlen, uint8_t* q, int tlen, uint8_t* t, const int8_t* mat, int od, int ed, int oi,
                        int ei, int xtra, int16_t* work, uint64_t* bbuf, unsigned long long* cells, bool* agree) {
  B2Kswr res[32];
  unsigned long long cl[32] = {0};
  W.run(G, [&](int lane) {
    EmuG g; g.w = &W; g.lane = lane; g.size = G; g.mask = 0;
    res[lane] = b2_align2_g(g, qlen, q, tlen, t, mat, od, ed, oi, ei, xtra, work, bbuf, &cl[lane]);
  });
  *agree = true;
  for (int l = 1; l < G; ++l) if (memcmp(&res[l], &res[0], sizeof(B2Kswr))) *agree = false;
  *cells = cl[0];
  return res[0];
}

int main(int argc, char** argv) {
  int trials = argc > 1 ? atoi(argv[1]) : 20000;
  srand48(argc > 2 ? atol(argv[2]) : 12345);
  long bad = 0, bad_g = 0, n_g = 0;
  EmuWarp W;
  for (int t = 0; t < trials; ++t) {
    int qlen = 20 + lrand48() % 240, tlen = 30 + lrand48() % 700;
    int size = (lrand48() % 3 == 0) ? 2 : 1;
    int a = 1 + lrand48() % 3, b = 1 + lrand48() % 6;
    int o_del = 1 + lrand48() % 8, e_del = 1 + lrand48() % 3, o_ins = 1 + lrand48() % 8, e_ins = 1 + lrand48() % 3;
    const bool zero_open = t % 11 == 5;   // -O x,0: the lazy-F early exit becomes observable; only the literal loop is exact
    if (zero_open) o_ins = 0;
    if (t % 3 == 0) { a = 1; b = 4; o_del = o_ins = 6; e_del = e_ins = 1; }
    int8_t mat[25]; int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? a : -b; mat[k++] = -1; }
    for (int j = 0; j < 5; ++j) mat[k++] = -1;
    std::vector<uint8_t> q(qlen), tg(tlen);
    for (auto& c : tg) c = lrand48() % 4;
    if (lrand48() % 4 == 0) for (auto& c : q) c = lrand48() % 4;
    else {
      int st = lrand48() % (tlen > qlen ? tlen - qlen + 1 : 1), p = st;
      double sub = 0.02 * (1 + lrand48() % 5), ind = 0.005 * (lrand48() % 6);
      for (int i = 0; i < qlen; ++i) {
        double u = drand48();
        if (u < ind) { q[i] = lrand48() % 4; continue; }
        if (u < 2 * ind) p += 1 + lrand48() % 4;
        uint8_t c = p < tlen ? tg[p] : lrand48() % 4; ++p;
        if (drand48() < sub) c = (c + 1 + lrand48() % 3) & 3;
        if (drand48() < 0.003) c = 4;
        q[i] = c;
      }
    }
    int xtra = B2_KSW_XSUBO | B2_KSW_XSTART | (size == 1 ? B2_KSW_XBYTE : 0) | (19 * a);
    if (lrand48() % 5 == 0) xtra = (size == 1 ? B2_KSW_XBYTE : 0) | B2_KSW_XSTOP | (20 + lrand48() % 100);
    int lanes = size == 1 ? 16 : 8, npad = ((qlen + lanes - 1) / lanes) * lanes;
    std::vector<int16_t> w1(4 * npad + 32), w2(4 * npad + 32), w3(4 * npad + 32);
    std::vector<uint64_t> b1(tlen + 1), b2(tlen + 1);
    std::vector<uint8_t> q2 = q, t2 = tg;
    B2Kswr r1 = b2_align2(qlen, q.data(), tlen, tg.data(), mat, o_del, e_del, o_ins, e_ins, xtra | B2_KSW_XLITERAL, w1.data(), b1.data());
    B2Kswr r2 = b2_align2(qlen, q2.data(), tlen, t2.data(), mat, o_del, e_del, o_ins, e_ins, xtra, w2.data(), b2.data());
    if (!zero_open && memcmp(&r1, &r2, sizeof(r1))) ++bad;
    if (t % 4 == 0) {  // the generic-group form, group sizes 1..16, against the scalar form (incl. the cell count)
      const int G = 1 << (lrand48() % 5);
      std::vector<uint8_t> q3 = q, t3 = tg;
      std::vector<uint64_t> b3(tlen + 1);
      unsigned long long c0 = 0, c3 = 0;
      std::vector<uint8_t> q4 = q, t4 = tg;
      B2Kswr r0 = b2_align2(qlen, q4.data(), tlen, t4.data(), mat, o_del, e_del, o_ins, e_ins, xtra, w1.data(), b1.data(), &c0);
      bool agree;
      B2Kswr r3 = run_group(W, G, qlen, q3.data(), tlen, t3.data(), mat, o_del, e_del, o_ins, e_ins, xtra, w3.data(), b3.data(), &c3, &agree);
      ++n_g;
      if (!agree || memcmp(&r0, &r3, sizeof(r0)) || c0 != c3 || q3 != q || t3 != tg) ++bad_g;
    }
  }
  printf("trials %d mismatches %ld\n", trials, bad);
  printf("generic-group trials %ld mismatches %ld\n", n_g, bad_g);
  return bad != 0 || bad_g != 0;
}
