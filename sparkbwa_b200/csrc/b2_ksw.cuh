// Shared pieces of the dynamic-programming kernels: the {H, E} cell pair, the CIGAR helpers, the ksw_align2 flag
// bits and result record (bwa/ksw.h:14-19,29-34).  The DP itself lives in b2_coop.cuh / b2_coopreg.cuh (lane groups),
// b2_xext.cuh and b2_galn.cuh (one alignment per lane).  The plain scalar forms that the host checker build
// (tests/hostsim) and the fuzzers compare against are test infrastructure: tests/hostsim/b2_scalar_dp.cuh.
#pragma once
#include "b2_types.h"

struct B2EH { int32_t h, e; };

#define B2_MINUS_INF (-0x40000000)

B2_HD inline int b2_push_cigar(uint32_t* cigar, int n, int op, int len) {
  if (n == 0 || op != (int)(cigar[n - 1] & 0xf)) { cigar[n++] = (uint32_t)len << 4 | op; }
  else cigar[n - 1] += (uint32_t)len << 4;
  return n;
}

// ---------------------------------------------------------------------------------------------
#define B2_KSW_XBYTE 0x10000
#define B2_KSW_XSTOP 0x20000
#define B2_KSW_XSUBO 0x40000
#define B2_KSW_XSTART 0x80000
#define B2_KSW_XLITERAL 0x100000  // test hook: run the reference's lazy-F loop literally

struct B2Kswr { int score, te, qe, score2, te2, tb, qb; };

B2_HD inline void b2_revseq(int l, uint8_t* s) {
  for (int i = 0; i < l >> 1; ++i) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; }
}


#if defined(B2_HOSTSIM) || defined(B2_WITH_SCALAR_DP)
#include "b2_scalar_dp.cuh"   // tests/hostsim: never part of libbwa.so
#endif
