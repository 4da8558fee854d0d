// Plain-data types shared by the host driver and the sm_100a kernels.
// Field meanings follow the reference's parameter/record blocks so that parity can be audited:
//   B2Opt      <- mem_opt_t      (bwa/bwamem.h:23-55, defaults bwa/bwamem.c:48-84)
//   B2Intv     <- bwtintv_t      (bwa/bwt.h:60-62)
//   B2Seed     <- mem_seed_t     (bwa/bwamem.c:168-172)
//   B2Chain    <- mem_chain_t    (bwa/bwamem.c:174-180)
//   B2Reg      <- mem_alnreg_t   (bwa/bwamem.h:57-75)
//   B2Pes      <- mem_pestat_t   (bwa/bwamem.h:79-83)
//   B2Aln      <- mem_aln_t      (bwa/bwamem.h:85-95)
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__
#define B2_D __device__
// Build experiment (-DB2_OUTLINE): keep the big leaf routines out of line on the device.  The finalisation kernel
// inlines to 123 k instructions and waits for instruction fetch half of the time (ncu stall_no_inst 49 %), but the
// out-of-line build (43 k instructions) measured SLOWER on the B200 (k_final_pe 10.3 -> 15.7 ms: by-reference
// arguments force the kernel parameter block into local memory), so inlining stays the default.
#if defined(B2_OUTLINE)
#define B2_NOINLINE __noinline__
#else
#define B2_NOINLINE
#endif
#else
#define B2_HD
#define B2_D
#define B2_NOINLINE
#endif

#define B2_F_PE 0x2
#define B2_F_NOPAIRING 0x4
#define B2_F_ALL 0x8
#define B2_F_NO_MULTI 0x10
#define B2_F_NO_RESCUE 0x20
#define B2_F_REF_HDR 0x100
#define B2_F_SOFTCLIP 0x200
#define B2_F_SMARTPE 0x400

struct B2Opt {
  int a, b;
  int o_del, e_del, o_ins, e_ins;
  int pen_unpaired;
  int pen_clip5, pen_clip3;
  int w;
  int zdrop;
  uint64_t max_mem_intv;
  int T;
  int flag;
  int min_seed_len;
  int min_chain_weight;
  int max_chain_extend;
  float split_factor;
  int split_width;
  int max_occ;
  int max_chain_gap;
  int n_threads;
  int chunk_size;
  float mask_level;
  float drop_ratio;
  float XA_drop_ratio;
  float mask_level_redun;
  float mapQ_coef_len;
  int mapQ_coef_fac;
  int max_ins;
  int max_matesw;
  int max_XA_hits, max_XA_hits_alt;
  int8_t mat[25];
};

struct B2Intv {
  uint64_t x[3];
  uint64_t info;
};

struct B2Seed {
  int64_t rbeg;
  int32_t qbeg, len;
  int32_t score;
  int32_t next;  // linked list while chaining (insertion order inside one chain)
};

struct B2Chain {
  int n, first, rid;
  uint32_t w;      // 29 bits used
  int8_t kept, is_alt;
  float frac_rep;
  int64_t pos;
  int seed_head, seed_tail;  // while chaining: linked list; after compaction: seed_head = offset
};

struct B2Reg {
  int64_t rb, re;
  int qb, qe;
  int rid;
  int score;
  int truesc;
  int sub;
  int alt_sc;
  int csub;
  int sub_n;
  int w;
  int seedcov;
  int secondary;
  int secondary_all;
  int seedlen0;
  int n_comp;
  int is_alt;
  float frac_rep;
  uint64_t hash;
};

// A plain 4-line FASTQ record inside a memory-mapped input file, as found by the reader's scan: position of its '@',
// sequence length, header-line length (after '@'), name length (up to the first white space), offset of the quality
// line relative to the '@'.  The device parses / trims / encodes reads from these (k_ingest_*).
struct B2Rec { uint64_t off; uint32_t l_seq, l_hdr, l_name, qual_rel; };

struct B2Pes {
  int low, high;
  int failed;
  double avg, std;
};

// Contig table entry on the device (bntann1_t without the strings; name bytes live in a blob).
struct B2Ann {
  int64_t offset;
  int32_t len;
  int32_t is_alt;
  int32_t name_off, name_len;
  int32_t anno_off, anno_len;
};

// Device-resident FM-index image (bwt_t + bntseq_t + pac).  `bwt` is the .bwt payload verbatim:
// 64-byte blocks {4 x u64 cumulative counts, 8 x u32 = 128 two-bit symbols} (bwa/bwt.h:35-37,72-78).
struct B2Index {
  const uint32_t* bwt;
  const uint64_t* sa;
  const uint8_t* pac;
  const B2Ann* anns;
  const char* names;  // contig name/anno blob
  uint64_t primary;
  uint64_t L2[5];
  uint64_t seq_len;
  uint64_t bwt_size;  // in u32 words
  uint64_t n_sa;
  int sa_intv;
  int sa_shift;       // log2(sa_intv)
  int64_t l_pac;
  int n_seqs;
  int max_name_len;
};

struct B2Aln {
  int64_t pos;
  int rid;
  int flag;
  uint32_t is_rev, is_alt, mapq, NM;
  int n_cigar;
  uint32_t* cigar;
  char* md;    // NUL-terminated MD string (the reference stores it right after the cigar)
  char* XA;    // NUL-terminated or null
  int score, sub, alt_sc;
};
