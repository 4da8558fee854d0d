// Equivalence fuzz: striped local SW with the max-plus carry scan (what the engine runs) vs the reference's
// lazy-F loop emulated literally (B2_KSW_XLITERAL), over random scorings, byte and word mode, related and
// unrelated sequences, ambiguous bases, XSUBO/XSTART and XSTOP modes.  Prints "mismatches N".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "b2_ksw.cuh"
int main(int argc, char** argv) {
  int trials = argc > 1 ? atoi(argv[1]) : 20000;
  srand48(argc > 2 ? atol(argv[2]) : 12345);
  long bad = 0;
  for (int t = 0; t < trials; ++t) {
    int qlen = 20 + lrand48() % 240, tlen = 30 + lrand48() % 700;
    int size = (lrand48() % 3 == 0) ? 2 : 1;
    int a = 1 + lrand48() % 3, b = 1 + lrand48() % 6;
    int o_del = 1 + lrand48() % 8, e_del = 1 + lrand48() % 3, o_ins = 1 + lrand48() % 8, e_ins = 1 + lrand48() % 3;
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
    std::vector<int16_t> w1(4 * npad), w2(4 * npad);
    std::vector<uint64_t> b1(tlen + 1), b2(tlen + 1);
    std::vector<uint8_t> q2 = q, t2 = tg;
    B2Kswr r1 = b2_align2(qlen, q.data(), tlen, tg.data(), mat, o_del, e_del, o_ins, e_ins, xtra | B2_KSW_XLITERAL, w1.data(), b1.data());
    B2Kswr r2 = b2_align2(qlen, q2.data(), tlen, t2.data(), mat, o_del, e_del, o_ins, e_ins, xtra, w2.data(), b2.data());
    if (memcmp(&r1, &r2, sizeof(r1))) ++bad;
  }
  printf("trials %d mismatches %ld\n", trials, bad);
  return bad != 0;
}
