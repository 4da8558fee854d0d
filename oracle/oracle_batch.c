/* TEST INFRASTRUCTURE ONLY (oracle side).  Calls the reference's own mem_process_seqs (bwa/bwamem.c:1172) on ONE
 * batch with a chosen n_processed, which the `bwa mem` command line cannot do: the absolute read index feeds the hash
 * tie-breaks (bwa/bwamem.c:527,1162,1167) and, as a 32-bit `id<<8`, mem_pair (bwa/bwamem_pair.c:222) -- it wraps for
 * pair indices >= 2^23, which a 20 M-pair job reaches.  Linked against oracle/_ref/libbwa.a, headers
 * taken from the reference tree at build time; prints the SAM records (no header) to stdout.
 *   oracle_batch <idxbase> <n_processed> <in1.fq> [in2.fq]                                                        */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "bwamem.h"
#include "kseq.h"
KSEQ_DECLARE(gzFile)

int main(int argc, char* argv[]) {
  if (argc < 4) { fprintf(stderr, "usage: oracle_batch <idxbase> <n_processed> <in1.fq> [in2.fq]\n"); return 1; }
  bwaidx_t* idx = bwa_idx_load(argv[1], BWA_IDX_ALL);
  if (!idx) return 1;
  int64_t n_processed = atoll(argv[2]);
  gzFile f1 = gzopen(argv[3], "r"), f2 = argc > 4 ? gzopen(argv[4], "r") : 0;
  if (!f1 || (argc > 4 && !f2)) return 1;
  kseq_t* k1 = kseq_init(f1);
  kseq_t* k2 = f2 ? kseq_init(f2) : 0;
  mem_opt_t* opt = mem_opt_init();
  if (k2) opt->flag |= MEM_F_PE;
  bwa_verbose = 1;
  int n = 0, i;
  bseq1_t* seqs = bseq_read(0x7fffffff, &n, k1, k2);   /* the whole input as one batch */
  mem_process_seqs(opt, idx->bwt, idx->bns, idx->pac, n_processed, n, seqs, 0);
  for (i = 0; i < n; ++i) if (seqs[i].sam) fputs(seqs[i].sam, stdout);
  return 0;
}
