// Unstable introsort whose tie order is observable in the reference's output.
//
// bwa/ksort.h:176-226 sorts with a median-of-three quicksort that leaves sub-ranges of <= 16
// elements for one final insertion sort, and falls back to a comb sort (:154-175) when the
// recursion depth budget (2*ceil(log2 n)) is exhausted.  With non-total orders (chains by weight
// bwamem.c:324, regions by end :391, regions by score/rb/qb :394) the permutation it produces
// decides which of several equal elements survives later filters, so the device code has to
// reproduce the same sequence of comparisons and swaps.  This is a from-scratch restatement of
// that procedure over indices instead of pointers.
#pragma once
#include "b2_types.h"

template <class T, class LT>
B2_HD inline void b2_insertsort(T* a, long s, long t, LT lt) {  // [s, t)
  for (long i = s + 1; i < t; ++i)
    for (long j = i; j > s && lt(a[j], a[j - 1]); --j) { T tmp = a[j]; a[j] = a[j - 1]; a[j - 1] = tmp; }
}

template <class T, class LT>
B2_HD inline void b2_combsort(T* a, long n, LT lt) {
  const double shrink = 1.2473309501039786540366528676643;
  long gap = n;
  int swapped;
  do {
    if (gap > 2) {
      gap = (long)((double)gap / shrink);
      if (gap == 9 || gap == 10) gap = 11;
    }
    swapped = 0;
    for (long i = 0; i < n - gap; ++i) {
      long j = i + gap;
      if (lt(a[j], a[i])) { T tmp = a[i]; a[i] = a[j]; a[j] = tmp; swapped = 1; }
    }
  } while (swapped || gap > 2);
  if (gap != 1) b2_insertsort(a, 0, n, lt);
}

template <class T, class LT>
B2_HD inline void b2_introsort(T* a, long n, LT lt) {
  if (n < 1) return;
  if (n == 2) {
    if (lt(a[1], a[0])) { T tmp = a[0]; a[0] = a[1]; a[1] = tmp; }
    return;
  }
  int d;
  for (d = 2; (1ul << d) < (unsigned long)n; ++d) {}
  struct Frame { long l, r; int depth; };
  Frame stack[64];  // only the larger side is ever deferred, so live frames <= log2(n)
  int top = 0;
  long s = 0, t = n - 1;
  d <<= 1;
  for (;;) {
    if (s < t) {
      if (--d == 0) {
        b2_combsort(a + s, t - s + 1, lt);
        t = s;
        continue;
      }
      long i = s, j = t, k = i + ((j - i) >> 1) + 1;
      if (lt(a[k], a[i])) {
        if (lt(a[k], a[j])) k = j;
      } else k = lt(a[j], a[i]) ? i : j;
      T rp = a[k];
      if (k != t) { T tmp = a[k]; a[k] = a[t]; a[t] = tmp; }
      for (;;) {
        do ++i; while (lt(a[i], rp));
        do --j; while (i <= j && lt(rp, a[j]));
        if (j <= i) break;
        T tmp = a[i]; a[i] = a[j]; a[j] = tmp;
      }
      { T tmp = a[i]; a[i] = a[t]; a[t] = tmp; }
      if (i - s > t - i) {
        if (i - s > 16) { stack[top].l = s; stack[top].r = i - 1; stack[top].depth = d; ++top; }
        s = t - i > 16 ? i + 1 : t;
      } else {
        if (t - i > 16) { stack[top].l = i + 1; stack[top].r = t; stack[top].depth = d; ++top; }
        t = i - s > 16 ? i - 1 : s;
      }
    } else {
      if (top == 0) {
        b2_insertsort(a, 0, n, lt);
        return;
      }
      --top;
      s = stack[top].l; t = stack[top].r; d = stack[top].depth;
    }
  }
}
