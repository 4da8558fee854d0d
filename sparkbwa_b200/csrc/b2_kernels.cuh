// Kernel bodies of the batch pipeline (mem_process_seqs, bwa/bwamem.c:1172-1201, re-staged for a GPU):
//   k_seed    SMEM seeding, one read per lane (state machine)                (mem_collect_intv)
//   k_sa      one task per sampled suffix-array position                     (bwt_sa + bns_intv2rid)
//   k_chain   chaining + chain filter, one read per lane                     (mem_chain, mem_chain_flt)
//   k_extend  seed extension + de-dup, one read per lane                     (mem_chain2aln, mem_sort_dedup_patch)
//   k_pestat  insert-size candidates, one pair per lane                      (mem_pestat, first loop)
//   k_final_* primary marking, rescue, pairing, CIGAR, SAM text              (worker2)
//   k_scan / k_gather  ordered compaction of the per-read SAM slots
// Every kernel is idempotent (inputs are never modified) so the host can re-run a stage with larger
// scratch when a capacity flag is raised.
#pragma once
#include "b2_backend.h"
#include "b2_types.h"
#include "b2_bwt.cuh"
#include "b2_chain.cuh"
#include "b2_align.cuh"
#include "b2_final.cuh"

#define B2_FLAG_OVF_INTV 1
#define B2_FLAG_OVF_SCRATCH 2
#define B2_FLAG_OVF_BIGN 32     // more tasks needed a slot of the outlier pool than it has
#define B2_FLAG_OVF_SLOT 4
#define B2_FLAG_ERR_TABLE 8
#define B2_FLAG_OVF_REG 16

enum { B2_CNT_OCC_BLOCKS = 0, B2_CNT_SA_CALLS, B2_CNT_SA_STEPS, B2_CNT_EXT_CELLS, B2_CNT_EXT_CALLS, B2_CNT_GLOB_CELLS,
       B2_CNT_GLOB_CALLS, B2_CNT_SW_CELLS, B2_CNT_SW_CALLS,
       B2_CNT_XEXT_CELLS,    // cells of k_xext_run alone      } the three DP kernels proper, for their own GCUPS figures
       B2_CNT_RESCSW_CELLS,  // cells of k_resc_sw alone       } (the category counters above include every in-kernel DP)
       B2_CNT_GALN_CELLS,    // cells of k_galn* alone         }
       B2_CNT_N };

struct B2Batch {  // device-resident read batch
  int n_reads;
  const uint8_t* seq;       // base codes 0..4, concatenated
  const char* qual;         // same offsets as seq
  const uint8_t* has_qual;  // per read
  const char* names;
  const char* comments;
  const int64_t* seq_off;
  const int32_t* l_seq;
  const int64_t* name_off;
  const int32_t* l_name;
  const int64_t* com_off;
  const int32_t* l_com;     // -1: no comment
};

struct B2Ctx {
  B2Opt opt;
  B2Index ix;
  B2Batch b;
  B2Tables tb;
  B2Fmt fmt;
  B2Pes pes[4];
  int64_t n_processed;
  int max_len;            // longest read in the batch
  int* flags;
  unsigned long long* counters;  // [B2_CNT_*], algorithmic work
  int* work;                     // dynamic work-distribution cursor of the running stage
  long long* dbg;                // optional per-item profile: [4*i] = {cycles, sw calls, glob calls, n regs}
  int32_t* dbg_seed;             // optional (B200BWA_DEBUG_TIMES): extension steps of every read in k_seed
  // seeding
  B2Intv* intv; int cap_intv; int32_t* n_intv; int32_t* n_tasks;
  B2Intv* seed_scratch;   // host checker build: per lane max_len+1 stack entries
  uint32_t* seed_spill;   // device: per lane 5*(max_len+1) words for what does not fit the shared-memory part
  int seed_E, seed_QW, seed_RS;  // per lane in shared memory: stack entries, words of 4-bit base codes, re-seed candidates
  // SA / chaining
  const int64_t* seed_off; int64_t n_seed_tasks;
  int32_t* sa_task_r; uint32_t* sa_task_iv;   // per lookup: read, interval << 16 | rank (k_sa_plan); null: k_sa searches
  B2Seed* raw; int32_t* raw_rid;
  B2Chain* chains; B2Seed* cseeds; int32_t* n_chain; int32_t* reg_cap;
  // extension
  const int64_t* reg_off; B2Reg* regs; int32_t* n_regs;
  const int64_t* chain_off; int64_t n_chain_tasks;   // (read, chain) tasks of the speculative extension kernel
  B2Reg* cand; uint8_t* cand_ok;                     // per seed: region computed ahead of the serial pass
  // per-lane scratch for the lane-per-read kernels
  uint8_t* scratch; size_t scratch_per_lane;
  // Outlier pool: big_n slots of big_slot bytes.  The per-lane regions above are sized for the ordinary read / pair; the
  // rare task that needs more (a read with thousands of seeds, a pair with dozens of regions, a wide traceback) takes a
  // slot (big_count: slots handed out in this launch) and runs again there, instead of every lane's region growing.
  uint8_t* big_pool; size_t big_slot; int big_n; int* big_count;
  // pestat
  int8_t* pe_dir; int64_t* pe_is;
  // planned mate-rescue attempts
  int32_t* resc_cnt; const int64_t* resc_off; B2RescTask* resc_tasks; B2Kswr* resc_res; int64_t n_resc;
  // seed extensions computed ahead of the chain-extension kernel (b2_xext.cuh)
  B2XextTask* xext_tasks; B2XextRes* xext_res;  // one task per (read, chain) and side; two result slots per seed
  int xext_qcap, xext_tcap;                     // staging capacity per lane (bases)
  int xext_cell8;                               // 1: {h, e} cells of 8 + 8 bits (scores of this batch stay below 256)
  int32_t *xext_key, *xext_hist, *xext_cur, *xext_perm;  // counting sort of the tasks by query length (longest first):
  const int64_t* xext_hoff;                              // the lanes of a warp then run extensions of similar size
  // global alignments computed ahead of the finalisation kernel (b2_galn.cuh)
  int32_t* galn_cnt; const int64_t* galn_off;   // [class][read] counts and their exclusive scans (row strides n+1, n+2)
  int64_t galn_base[B2_GALN_CLASSES], galn_n[B2_GALN_CLASSES];
  int galn_blk[B2_GALN_CLASSES + 1];            // first block of each class in the combined k_galn launch
  B2GalnTask* galn_tasks; B2GalnRes* galn_res; int32_t* galn_table;
  B2GalnCache galn;
  // finalisation in two launches: pairs with many regions first, a warp each (k_final_plan)
  uint8_t* fin_heavy; int32_t* fin_list; int fin_heavy_min, fin_mode;
  int fin_fast;           // bytes of shared memory per task group of the pair finalisation (B2Scratch::alloc_fast)
  // SAM
  char* slots; int slot_size; int32_t* sam_len; const int64_t* sam_off; char* sam_out;
};

B2_HD inline B2Read b2_make_read(const B2Batch& b, int r, int copy_comment) {
  B2Read s;
  s.seq = b.seq + b.seq_off[r];
  s.qual = b.has_qual[r] ? b.qual + b.seq_off[r] : 0;
  s.name = b.names + b.name_off[r];
  s.l_seq = b.l_seq[r]; s.l_name = b.l_name[r];
  if (copy_comment && b.l_com[r] >= 0) { s.comment = b.comments + b.com_off[r]; s.l_comment = b.l_com[r]; }
  else { s.comment = 0; s.l_comment = 0; }
  return s;
}

// One task per lane at a time; lanes pull tasks from the stage's work cursor.  Each trip of the outer loop performs
// ONE bidirectional extension for every lane that has a request pending (convergent, expensive: two 64-byte block
// fetches + popcounts) and then the lane-local bookkeeping of the search (divergent, cheap, shared memory only)
// -- see B2SeedFsm.  MODE 0 = SMEM passes 1+2 of a read, MODE 1 = its forward-only pass 3; two kernels because the
// second needs no interval stack and far fewer registers, so many more of its lanes fit an SM.
// Shared memory per block: blockDim.x * (3*E + QW + 2*RS) words (MODE 1: E = RS = 0).
template <int MODE>
B2_D inline void b2_seed_body(const B2Ctx& c) {
#if defined(B2_HOSTSIM)
  B2OccScalar occ;
  B2SeedFsm<B2SeedMemHost, MODE> m;
  m.mem.stk = c.seed_scratch + (size_t)B2_GTID() * (size_t)(c.max_len + 1);
  m.mem.rsv = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (size_t)(c.max_len + 1));
  B2Intv* chk = (B2Intv*)malloc(sizeof(B2Intv) * (size_t)(2 * c.cap_intv + 3 * (c.max_len + 1)));
#else
  extern __shared__ uint32_t b2_seed_smem[];
  B2OccLane occ;
  B2SeedFsm<B2SeedMemDev, MODE> m;
  {
    const int T = blockDim.x;
    const int E = MODE ? 0 : c.seed_E, RS = MODE ? 0 : c.seed_RS;
    m.mem.stride = T; m.mem.E = E; m.mem.QW = c.seed_QW; m.mem.RS = RS;
    m.mem.s = b2_seed_smem + threadIdx.x;
    m.mem.qs = b2_seed_smem + (size_t)3 * E * T + threadIdx.x;
    m.mem.rs = b2_seed_smem + (size_t)(3 * E + c.seed_QW) * T + threadIdx.x;
    m.mem.g = c.seed_spill + (size_t)B2_GTID() * 5 * (size_t)(c.max_len + 1);
    m.mem.grs = m.mem.g + 3 * (size_t)(c.max_len + 1);
  }
#endif
  m.phase = B2S_IDLE;
  int r = -1, need = 0, trips = 0;
  bool alive = true;
  int32_t* n_out = MODE ? c.n_tasks : c.n_intv;  // (n_tasks holds the pass-3 count until k_seed_fin)
  for (;;) {
    while (alive && !need) {  // next task of this lane
#if defined(__CUDA_ARCH__)
      r = atomicAdd(c.work + MODE, 1);
#else
      r = c.work[MODE]++;
#endif
      if (r >= c.b.n_reads) { alive = false; break; }
      const int len = c.b.l_seq[r];
      n_out[r] = 0;
      if (len < c.opt.min_seed_len) continue;
      need = m.begin(c.opt, c.ix, c.b.seq + c.b.seq_off[r], len, c.intv + (size_t)r * c.cap_intv, c.cap_intv);
      trips = 0;
    }
    // the warp-wide vote is also the reconvergence point: every lane with a pending request runs the extension
    // below in the same pass
#if defined(__CUDA_ARCH__)
    if (__ballot_sync(0xffffffffu, need) == 0) break;
#else
    if (!need) break;
#endif
    if (need) {
      B2Intv okc;
      okc.info = 0;
      occ.child(c.ix, m.p, m.rc, m.rback, okc);
#if defined(B2_HOSTSIM)
      static const bool dump_steps = getenv("B200BWA_DEBUG_SEED_STEPS") != 0;
      static long steps[4]; static int steps_read = -1;
      if (dump_steps) { if (steps_read != r) { steps_read = r; steps[0] = steps[1] = steps[2] = steps[3] = 0; } ++steps[m.phase]; }
#endif
      ++trips;
      need = m.consume(c.opt, c.ix, okc);
#if defined(B2_HOSTSIM)
      if (dump_steps && !need) fprintf(stderr, "[D::seed_steps] %d fwd %ld bwd %ld s3 %ld n %d\n", r, steps[B2S_FWD], steps[B2S_BWD], steps[B2S_S3], m.n);
#endif
      if (!need) {  // task finished; k_seed_fin joins the two runs, sorts the intervals and counts the SA tasks
        int n = m.n;
        if (n > c.cap_intv) { B2_ATOMIC_OR(c.flags, B2_FLAG_OVF_INTV); n = c.cap_intv; }
        n_out[r] = n;
        if (MODE == 0 && c.dbg_seed) c.dbg_seed[r] = trips;
#if defined(B2_HOSTSIM)
        if (MODE && c.n_intv[r] + n <= c.cap_intv) {  // cross-check against the plain nested-loop form (x[1] aside)
          B2OccScalar occ2;
          B2Intv* mine = chk + c.cap_intv + 3 * (c.max_len + 1);
          const int n12 = c.n_intv[r];
          memcpy(mine, m.out, sizeof(B2Intv) * (size_t)n12);
          for (int k = 0; k < n; ++k) mine[n12 + k] = m.out[c.cap_intv - 1 - k];
          n += n12;
          b2_sort_intv(mine, n);
          int n2 = b2_collect_intv(c.opt, c.ix, occ2, m.len, c.b.seq + c.b.seq_off[r], chk, c.cap_intv, chk + c.cap_intv);
          int bad = n2 <= c.cap_intv && n2 != n;
          for (int k = 0; !bad && k < n && n2 <= c.cap_intv; ++k)
            bad = chk[k].x[0] != mine[k].x[0] || chk[k].x[2] != mine[k].x[2] || chk[k].info != mine[k].info ||
                  (mine[k].x[1] != 0 && mine[k].x[1] != chk[k].x[1]);
          if (bad) {
            fprintf(stderr, "[hostsim] seeding state machine differs from the nested form on read %d (%d vs %d intervals)\n", r, n, n2);
            abort();
          }
        }
#endif
      }
    }
  }
  if (occ.n_blk) B2_ATOMIC_ADD(&c.counters[B2_CNT_OCC_BLOCKS], occ.n_blk);
#if defined(B2_HOSTSIM)
  free(chk); free(m.mem.rsv);
#endif
}

B2_KERNEL k_seed(B2Ctx c) { b2_seed_body<0>(c); }
B2_KERNEL k_seed3(B2Ctx c) { b2_seed_body<1>(c); }

// per read: intervals sorted by (start, end) as mem_collect_intv leaves them (bwamem.c:161) + number of SA lookups
B2_KERNEL k_seed_fin(B2Ctx c) {
  for (int r = B2_GTID(); r < c.b.n_reads; r += B2_GSIZE()) {
    B2Intv* out = c.intv + (size_t)r * c.cap_intv;
    const int n12 = c.n_intv[r], n3 = c.n_tasks[r];
    if (n12 + n3 > c.cap_intv) { B2_ATOMIC_OR(c.flags, B2_FLAG_OVF_INTV); c.n_tasks[r] = 0; continue; }
    const int gap = c.cap_intv - n3 - n12;  // close the gap between the two runs (their order is irrelevant: sorted below)
    for (int k = 0; k < n3 && k < gap; ++k) out[n12 + k] = out[c.cap_intv - 1 - k];
    const int n = n12 + n3;
    c.n_intv[r] = n;
    b2_sort_intv(out, n);
    long tasks = 0;
    for (int i = 0; i < n; ++i) tasks += out[i].x[2] > (uint64_t)c.opt.max_occ ? c.opt.max_occ : (long)out[i].x[2];
    c.n_tasks[r] = (int32_t)tasks;
  }
}

// (read, interval, rank inside the interval) of every suffix-array lookup, written by the read's own lane once the
// offsets are known: k_sa then fetches a lookup's description with two coalesced loads where a search for the read
// (17 dependent probes of seed_off) and a scan of its intervals would sit on the path of the lane's refill -- and a
// refill is executed by the few lanes of a warp that finished a walk in the same trip.
B2_KERNEL k_sa_plan(B2Ctx c) {
  for (int r = B2_GTID(); r < c.b.n_reads; r += B2_GSIZE()) {
    int64_t t = c.seed_off[r];
    const B2Intv* iv = c.intv + (size_t)r * c.cap_intv;
    const int n = c.n_intv[r];
    for (int i = 0; i < n; ++i) {
      const int cnt = iv[i].x[2] > (uint64_t)c.opt.max_occ ? c.opt.max_occ : (int)iv[i].x[2];
      for (int idx = 0; idx < cnt; ++idx, ++t) { c.sa_task_r[t] = r; c.sa_task_iv[t] = (uint32_t)i << 16 | (uint32_t)idx; }
    }
  }
}

// Form without k_sa_plan's tables (max_occ or the interval capacity beyond 16 bits): the lane finds its lookup's read by
// searching seed_off.
B2_KERNEL k_sa_search(B2Ctx c) {
  // One suffix-array lookup per lane at a time: an LF walk to the next sampled row, geometric in length (mean 31 steps,
  // the longest of 32 about four times that).  A lane that finishes takes the next lookup from the stage's cursor right
  // away instead of idling until the warp's longest walk ends; every trip of the loop is ONE LF step for all lanes
  // that hold a lookup.
  unsigned long long steps_total = 0, calls = 0;
  const uint64_t mask = (uint64_t)c.ix.sa_intv - 1;
  int64_t t = 0;
  uint64_t k = 0;
  unsigned steps = 0;
  int32_t qbeg = 0, slen = 0;
  bool have = false, alive = true;
  for (;;) {
    if (!have && alive) {
#if defined(__CUDA_ARCH__)
      t = (int64_t)atomicAdd(c.work + 2, 1);
#else
      t = (int64_t)c.work[2]++;
#endif
      if (t >= c.n_seed_tasks) alive = false;
      else {
        int r, i;
        long idx;
        if (c.sa_task_r) {   // planned (k_sa_plan)
          r = c.sa_task_r[t];
          const uint32_t pk = c.sa_task_iv[t];
          i = (int)(pk >> 16); idx = (long)(pk & 0xffff);
        } else {
          int lo = 0, hi = c.b.n_reads;  // last r with seed_off[r] <= t
          while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (c.seed_off[mid] <= t) lo = mid; else hi = mid; }
          r = lo;
          idx = (long)(t - c.seed_off[r]);
          const B2Intv* iv0 = c.intv + (size_t)r * c.cap_intv;
          const int n = c.n_intv[r];
          for (i = 0; i < n; ++i) {
            long cnt = iv0[i].x[2] > (uint64_t)c.opt.max_occ ? c.opt.max_occ : (long)iv0[i].x[2];
            if (idx < cnt) break;
            idx -= cnt;
          }
        }
        const B2Intv* iv = c.intv + (size_t)r * c.cap_intv;
        const B2Intv& p = iv[i];
        const long step = p.x[2] > (uint64_t)c.opt.max_occ ? (long)(p.x[2] / c.opt.max_occ) : 1;
        slen = (int32_t)((uint32_t)p.info - (uint32_t)(p.info >> 32));
        qbeg = (int32_t)(p.info >> 32);
        k = p.x[0] + (uint64_t)(idx * step);
        steps = 0;
        have = true;
      }
    }
#if defined(__CUDA_ARCH__)
    if (__ballot_sync(0xffffffffu, have) == 0) break;
#else
    if (!have) break;
#endif
    if (have) {
#if defined(__CUDA_ARCH__)
      if (k & mask) { ++steps; k = b2_inv_psi_lane(c.ix, k); }
#else
      if (k & mask) { ++steps; k = b2_inv_psi(c.ix, k); }
#endif
      else {
        B2Seed s;
        s.rbeg = (int64_t)(steps + c.ix.sa[k >> c.ix.sa_shift]);
        s.qbeg = qbeg;
        s.len = s.score = slen;
        s.next = -1;
        c.raw[t] = s;   // (its contig id: k_sa_rid, all lanes together)
        steps_total += steps; ++calls;
        have = false;
      }
    }
  }
  if (calls) { B2_ATOMIC_ADD(&c.counters[B2_CNT_SA_CALLS], calls); B2_ATOMIC_ADD(&c.counters[B2_CNT_SA_STEPS], steps_total); }
}

// One suffix-array lookup per lane at a time: an LF walk to the next sampled row, geometric in length (mean 31 steps, the
// longest of 32 about four times that); every trip of the loop is ONE LF step -- one dependent block fetch -- for all lanes
// that hold a lookup.  What surrounds a walk is kept off that path by two small per-lane pipelines that advance one stage
// per trip, each stage consuming what the previous trip requested:
//   next lookup:  index from the stage's cursor -> its description (k_sa_plan's tables) -> its interval -> ready;
//   finished one: the sampled SA entry is requested when the walk ends, the seed is written a trip later.
// A lane that finishes a walk therefore continues with the next one in the same trip, and no trip waits for more than one
// memory round trip (a refill done in place is a chain of four, executed by the few lanes that finished together).
B2_KERNEL k_sa(B2Ctx c) {
  unsigned long long steps_total = 0, calls = 0;
  const uint64_t mask = (uint64_t)c.ix.sa_intv - 1;
  int64_t t = 0;
  uint64_t k = 0;
  unsigned steps = 0;
  int32_t qbeg = 0, slen = 0;
  bool have = false;
  int pf = 0;   // next lookup: 0 nothing requested, 1 index, 2 description, 3 interval requested, 4 ready, 5 no lookups left
  int64_t pf_t = 0;
  int pf_r = 0;
  uint32_t pf_pk = 0;
  uint64_t pf_x0 = 0, pf_x2 = 0, pf_info = 0;
  bool fin = false;   // finished lookup whose seed is still to be written
  int64_t fin_t = 0;
  uint64_t fin_sa = 0;
  unsigned fin_steps = 0;
  int32_t fin_qbeg = 0, fin_slen = 0;
  for (;;) {
    if (fin) {
      B2Seed s;
      s.rbeg = (int64_t)(fin_steps + fin_sa);
      s.qbeg = fin_qbeg;
      s.len = s.score = fin_slen;
      s.next = -1;
      c.raw[fin_t] = s;   // (its contig id: k_sa_rid, all lanes together)
      fin = false;
    }
    if (pf == 3) pf = 4;
    else if (pf == 2) {
      const B2Intv* p = c.intv + (size_t)pf_r * c.cap_intv + (pf_pk >> 16);
      pf_x0 = p->x[0]; pf_x2 = p->x[2]; pf_info = p->info;
      pf = 3;
    } else if (pf == 1) {
      if (pf_t >= c.n_seed_tasks) pf = 5;
      else { pf_r = c.sa_task_r[pf_t]; pf_pk = c.sa_task_iv[pf_t]; pf = 2; }
    } else if (pf == 0) {
#if defined(__CUDA_ARCH__)
      pf_t = (int64_t)atomicAdd(c.work + 2, 1);
#else
      pf_t = (int64_t)c.work[2]++;
#endif
      pf = 1;
    }
    if (!have && pf == 4) {
      const long idx = (long)(pf_pk & 0xffff);
      const long step = pf_x2 > (uint64_t)c.opt.max_occ ? (long)(pf_x2 / c.opt.max_occ) : 1;
      slen = (int32_t)((uint32_t)pf_info - (uint32_t)(pf_info >> 32));
      qbeg = (int32_t)(pf_info >> 32);
      k = pf_x0 + (uint64_t)(idx * step);
      t = pf_t;
      steps = 0;
      have = true;
      pf = 0;
    }
#if defined(__CUDA_ARCH__)
    if (__ballot_sync(0xffffffffu, have || fin || pf != 5) == 0) break;
#else
    if (!(have || fin || pf != 5)) break;
#endif
    if (have) {
#if defined(__CUDA_ARCH__)
      if (k & mask) { ++steps; k = b2_inv_psi_lane(c.ix, k); }
#else
      if (k & mask) { ++steps; k = b2_inv_psi(c.ix, k); }
#endif
      else {
        fin_sa = c.ix.sa[k >> c.ix.sa_shift];
        fin_t = t; fin_steps = steps; fin_qbeg = qbeg; fin_slen = slen;
        fin = true;
        steps_total += steps; ++calls;
        have = false;
      }
    }
  }
  if (calls) { B2_ATOMIC_ADD(&c.counters[B2_CNT_SA_CALLS], calls); B2_ATOMIC_ADD(&c.counters[B2_CNT_SA_STEPS], steps_total); }
}

// contig of every seed (bns_intv2rid, bwa/bntseq.c:343-359; negative: the seed bridges two contigs or the strands and
// mem_chain skips it): two searches of the contig table per seed, kept out of k_sa's end-of-walk path
B2_KERNEL k_sa_rid(B2Ctx c) {
  for (int64_t t = B2_GTID(); t < c.n_seed_tasks; t += B2_GSIZE()) {
    const int64_t rb = c.raw[t].rbeg;
    c.raw_rid[t] = b2_intv2rid(c.ix, rb, rb + c.raw[t].len);
  }
}

B2_HD inline B2Scratch b2_lane_scratch(const B2Ctx& c, int lane) {
  B2Scratch sc;
  sc.base = c.scratch + (size_t)lane * c.scratch_per_lane;
  sc.cap = c.scratch_per_lane; sc.used = 0; sc.overflow = 0;
  return sc;
}

// scratch + lane-group identity of a task group (B2_TASK_GROUP lanes per read / pair)
B2_HD inline B2Scratch b2_group_scratch(const B2Ctx& c, int* gid, int* n_groups, const int G) {
  *gid = B2_GTID() / G;
  *n_groups = B2_GSIZE() / G;
  B2Scratch sc = b2_lane_scratch(c, *gid);
#if defined(__CUDA_ARCH__)
  sc.gsize = G;
  sc.lane = threadIdx.x & (G - 1);
  sc.gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1)) << ((threadIdx.x & 31) & ~(G - 1));
#endif
  return sc;
}

// Next task index for this group: groups pull reads/pairs from a shared cursor, so a group stuck on an
// expensive read (repeats: hundreds of chains, dozens of rescue windows) does not also own a fixed share of the rest.
B2_D inline int b2_next_task(const B2Ctx& c, const B2Scratch& sc) {
#if defined(__CUDA_ARCH__)
  int r = 0;
  if (sc.lane == 0) r = atomicAdd(c.work, 1);
  if (sc.gsize > 1) r = __shfl_sync(sc.gmask, r, 0, sc.gsize);
  return r;
#else
  return (*c.work)++;
#endif
}

// After a task body: true if it overflowed the lane's own region and now holds a slot of the outlier pool to run again in.
B2_D inline bool b2_retry_big(const B2Ctx& c, B2Scratch& sc) {
  if (!sc.overflow || sc.big || !c.big_pool) return false;
  int slot = 0;
#if defined(__CUDA_ARCH__)
  if (sc.lane == 0) slot = atomicAdd(c.big_count, 1);
  if (sc.gsize > 1) slot = __shfl_sync(sc.gmask, slot, 0, sc.gsize);
#else
  slot = (*c.big_count)++;
#endif
  if (slot >= c.big_n) return false;
  sc.small_base = sc.base; sc.small_cap = sc.cap;
  sc.base = c.big_pool + (size_t)slot * c.big_slot; sc.cap = c.big_slot;
  sc.big = 1; sc.overflow = 0;
  return true;
}
// End of a task: reports an overflow that remains (inside a slot: the slots are too small; outside: no slot was left) and
// returns to the lane's own region.  Returns 1 if the task's results are void.
B2_D inline int b2_end_task(const B2Ctx& c, B2Scratch& sc) {
  const int ovf = sc.overflow;
  if (ovf) B2_ATOMIC_OR(c.flags, (sc.big || !c.big_pool) ? B2_FLAG_OVF_SCRATCH : B2_FLAG_OVF_BIGN);
  if (sc.big) { sc.base = sc.small_base; sc.cap = sc.small_cap; sc.big = 0; }
  sc.overflow = 0;
  return ovf;
}

B2_D inline void b2_flush_counters(const B2Ctx& c, const B2Scratch& sc) {
  if (sc.lane != 0) return;
  if (sc.ext_calls) { B2_ATOMIC_ADD(&c.counters[B2_CNT_EXT_CELLS], sc.ext_cells); B2_ATOMIC_ADD(&c.counters[B2_CNT_EXT_CALLS], (unsigned long long)sc.ext_calls); }
  if (sc.glob_calls) { B2_ATOMIC_ADD(&c.counters[B2_CNT_GLOB_CELLS], sc.glob_cells); B2_ATOMIC_ADD(&c.counters[B2_CNT_GLOB_CALLS], (unsigned long long)sc.glob_calls); }
  if (sc.sw_calls) { B2_ATOMIC_ADD(&c.counters[B2_CNT_SW_CELLS], sc.sw_cells); B2_ATOMIC_ADD(&c.counters[B2_CNT_SW_CALLS], (unsigned long long)sc.sw_calls); }
}

B2_KERNEL k_chain(B2Ctx c) {
  B2Scratch sc = b2_lane_scratch(c, B2_GTID());
  for (int r = b2_next_task(c, sc); r < c.b.n_reads; r = b2_next_task(c, sc)) {
    const int64_t off = c.seed_off[r];
    const int n_raw = (int)(c.seed_off[r + 1] - off);
    int kept = 0, cap = 0;
    if (n_raw > 0) {
      do {
        sc.reset(0);
        kept = cap = 0;
        const int node_cap = n_raw / 3 + 4;
        B2Seed* raw = (B2Seed*)sc.alloc(sizeof(B2Seed) * (size_t)n_raw, 8);
        B2Chain* tree_ch = (B2Chain*)sc.alloc(sizeof(B2Chain) * (size_t)n_raw, 8);
        int* order = (int*)sc.alloc(sizeof(int) * (size_t)n_raw, 4);
        B2BtNode* nodes = (B2BtNode*)sc.alloc(sizeof(B2BtNode) * (size_t)node_cap, 4);
        if (!sc.overflow) {
          for (int i = 0; i < n_raw; ++i) { raw[i] = c.raw[off + i]; raw[i].next = c.raw_rid[off + i]; }
          int ovf = 0;
          kept = b2_chain_read(c.opt, c.ix.l_pac, c.ix.anns, c.b.l_seq[r], c.intv + (size_t)r * c.cap_intv, c.n_intv[r],
                               raw, n_raw, tree_ch, c.chains + off, c.cseeds + off, order, nodes, node_cap, &ovf);
          if (ovf) B2_ATOMIC_OR(c.flags, B2_FLAG_OVF_SCRATCH);
          for (int i = 0; i < kept; ++i) cap += c.chains[off + i].n;
        }
      } while (b2_retry_big(c, sc));
      if (b2_end_task(c, sc)) kept = cap = 0;
    }
    c.n_chain[r] = kept;
    c.reg_cap[r] = cap;
  }
}

// mem_flt_chained_seeds: one lane group per read (only launched when the batch holds reads long enough for the
// filter to act at all); compacts the seeds of every kept chain in place.
B2_KERNEL k_flt_seeds(B2Ctx c) {
  int gid, n_groups;
  B2Scratch sc = b2_group_scratch(c, &gid, &n_groups, B2_G_RESC);
  for (int r = b2_next_task(c, sc); r < c.b.n_reads; r = b2_next_task(c, sc)) {
    sc.gsync();
    sc.reset(0);
    const int l_seq = c.b.l_seq[r];
    int min_hsp = 0, err = 0;
    if (!b2_flt_seeds_threshold(c.opt, c.tb, l_seq, &min_hsp, &err)) { if (err) B2_ATOMIC_OR(c.flags, B2_FLAG_ERR_TABLE); continue; }
    if (err) B2_ATOMIC_OR(c.flags, B2_FLAG_ERR_TABLE);
    const int64_t off = c.seed_off[r];
    const uint8_t* query = c.b.seq + c.b.seq_off[r];
    int cap = 0;
    for (int i = 0; i < c.n_chain[r]; ++i) {
      B2Chain& ch = c.chains[off + i];
      B2Seed* seeds = c.cseeds + off + ch.seed_head;
      int k = 0;
      for (int j = 0; j < ch.n; ++j) {
        B2Seed s = seeds[j];
        int score = b2_seed_sw(c.opt, c.ix, l_seq, query, s, sc);
        if (sc.overflow) break;
        if (score < 0 || score >= min_hsp) {
          s.score = score < 0 ? s.len * c.opt.a : score;
          sc.gsync();
          if (sc.lane == 0) seeds[k] = s;
          sc.gsync();
          ++k;
        }
      }
      if (sc.overflow) break;
      sc.gsync();
      if (sc.lane == 0) ch.n = k;
      sc.gsync();
      cap += k;
    }
    if (b2_end_task(c, sc)) continue;   // (no second run in a pool slot here: the compaction is in place)
    if (sc.lane == 0) c.reg_cap[r] = cap;
  }
  b2_flush_counters(c, sc);
}

// ---- inter-task seed extension ahead of k_ext_chain (b2_xext.cuh) ----------------------------------------------
// side 0: the first band try of the LEFT extension of every chain's best seed (the seed mem_chain2aln extends
// first: largest score, ties to the later one); side 1: of its RIGHT extension, whose start score is the left
// result (so it is planned once the left pass has run, and only when that first try was final).
B2_KERNEL k_xext_plan(B2Ctx c, int side) {
  for (int64_t t = B2_GTID(); t < c.n_chain_tasks; t += B2_GSIZE()) {
    int lo = 0, hi = c.b.n_reads;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (c.chain_off[mid] <= t) lo = mid; else hi = mid; }
    const int r = lo;
    const int64_t off = c.seed_off[r];
    const B2Chain& ch = c.chains[off + (t - c.chain_off[r])];
    const B2Seed* seeds = c.cseeds + off + ch.seed_head;
    const int l_seq = c.b.l_seq[r];
    B2XextTask task;
    task.qlen = -1;
    if (ch.n > 0) {
      int si = 0;
      for (int i = 1; i < ch.n; ++i) if (seeds[i].score >= seeds[si].score) si = i;
      const B2Seed s = seeds[si];
      const int64_t gidx = off + ch.seed_head + si;
      int64_t rmax[2];
      b2_chain_window(c.opt, c.ix, l_seq, ch, seeds, rmax);
      if (side == 0) {
        if (s.qbeg > 0) {
          task.tpos = s.rbeg - 1; task.tdir = -1; task.tlen = (int)(s.rbeg - rmax[0]);
          task.qsrc = c.b.seq_off[r] + s.qbeg - 1; task.qdir = -1; task.qlen = s.qbeg;
          task.h0 = s.len * c.opt.a; task.w = c.opt.w; task.end_bonus = c.opt.pen_clip5; task.slot = (int32_t)(2 * gidx);
        }
      } else if (s.qbeg + s.len != l_seq) {
        int sc0 = s.len * c.opt.a;
        bool ok = true;
        if (s.qbeg > 0) {
          const B2XextRes L = c.xext_res[2 * gidx];
          ok = L.w == c.opt.w && L.h0 == sc0 && L.max_off < (c.opt.w >> 1) + (c.opt.w >> 2);
          sc0 = L.score;
        }
        if (ok) {
          const int qe = s.qbeg + s.len;
          task.tpos = s.rbeg + s.len; task.tdir = 1; task.tlen = (int)(rmax[1] - (s.rbeg + s.len));
          task.qsrc = c.b.seq_off[r] + qe; task.qdir = 1; task.qlen = l_seq - qe;
          task.h0 = sc0; task.w = c.opt.w; task.end_bonus = c.opt.pen_clip3; task.slot = (int32_t)(2 * gidx + 1);
        }
      }
      if (task.qlen >= 0 && (task.qlen < 1 || task.qlen > c.xext_qcap || task.tlen < 0 || task.tlen > c.xext_tcap ||
                             task.h0 + task.qlen * c.opt.a >= (c.xext_cell8 ? 255 : 65000) || 2 * gidx + 1 > 0x7fffffff))
        task.qlen = -1;
    }
    c.xext_tasks[t] = task;
    const int key = task.qlen < 0 ? c.xext_qcap : c.xext_qcap - task.qlen;  // invalid tasks go to the last bin
    c.xext_key[t] = key;
    B2_ATOMIC_ADD(&c.xext_hist[key], 1);
  }
}

B2_KERNEL k_xext_scatter(B2Ctx c) {
  for (int64_t t = B2_GTID(); t < c.n_chain_tasks; t += B2_GSIZE()) {
    const int key = c.xext_key[t];
    const int64_t pos = c.xext_hoff[key] + B2_ATOMIC_ADD(&c.xext_cur[key], 1);
    c.xext_perm[pos] = (int32_t)t;
  }
}

// one extension per lane; the warp's shared-memory block holds, lane-interleaved, eh[] and the two staged sequences
template <class Cell>
B2_D inline void b2_xext_run_body(const B2Ctx& c) {
  const int qw = c.xext_qcap / 8 + 1, tw = c.xext_tcap / 8 + 1;
  const int ehw = (int)(((size_t)(c.xext_qcap + 2) * sizeof(Cell) + 3) / 4);   // 32-bit words of one lane's eh[]
  B2XMemT<Cell> M;
#if defined(B2_HOSTSIM)
  M.stride = 1;
  uint32_t* hbase = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(ehw + qw + tw));
  M.eh = (Cell*)hbase;
  M.q = hbase + ehw; M.t = M.q + qw;
#else
  extern __shared__ uint32_t b2_xext_smem[];
  const int lane = threadIdx.x & 31;
  uint32_t* base = b2_xext_smem + (size_t)(threadIdx.x >> 5) * 32 * (size_t)(ehw + qw + tw);
  M.stride = 32;
  M.eh = (Cell*)base + lane; M.q = base + (size_t)32 * ehw + lane; M.t = M.q + (size_t)32 * qw;
#endif
  unsigned long long cells = 0, calls = 0;
  const int64_t n_valid = c.xext_hoff[c.xext_qcap];  // the invalid tasks sit in the last bin
  for (int64_t t = B2_GTID(); t < n_valid; t += B2_GSIZE()) {
    const B2XextTask task = c.xext_tasks[c.xext_perm[t]];
    const uint8_t* qs = c.b.seq + task.qsrc;
    const int qdir = task.qdir;
    b2_xext_stage(M.q, M.stride, task.qlen, [&](int k) { return (int)qs[(int64_t)k * qdir]; });
    const int64_t tpos = task.tpos, l_pac = c.ix.l_pac;
    const int tdir = task.tdir;
    const uint8_t* pac = c.ix.pac;
    b2_xext_stage(M.t, M.stride, task.tlen, [&](int k) { return b2_base_at(pac, l_pac, tpos + (int64_t)k * tdir); });
    B2XextRes res;
    res.w = task.w; res.h0 = task.h0;
    res.score = b2_xext_one(M, task.qlen, task.tlen, c.opt.mat, c.opt.o_del, c.opt.e_del, c.opt.o_ins, c.opt.e_ins, task.w,
                            task.end_bonus, c.opt.zdrop, task.h0, &res.qle, &res.tle, &res.gtle, &res.gscore, &res.max_off, &cells);
    ++calls;
    c.xext_res[task.slot] = res;
  }
  if (calls) {
    B2_ATOMIC_ADD(&c.counters[B2_CNT_EXT_CELLS], cells); B2_ATOMIC_ADD(&c.counters[B2_CNT_EXT_CALLS], calls);
    B2_ATOMIC_ADD(&c.counters[B2_CNT_XEXT_CELLS], cells);
  }
#if defined(B2_HOSTSIM)
  free(hbase);
#endif
}

B2_KERNEL k_xext_run(B2Ctx c) {
  if (c.xext_cell8) b2_xext_run_body<uint16_t>(c);
  else b2_xext_run_body<uint32_t>(c);
}

// One task group per (read, kept chain): extends the chain's seeds against a chain-local region list and
// leaves every region it computes in cand[] for the serial per-read pass below (mem_chain2aln's DP, hoisted
// out of the per-read control flow so that a read with hundreds of chains is spread over as many groups).
B2_KERNEL_LB(128, B2_LB_EXT) k_ext_chain(B2Ctx c) {
  int gid, n_groups;
  B2Scratch sc = b2_group_scratch(c, &gid, &n_groups, B2_G_EXT);
  for (int64_t t = b2_next_task(c, sc); t < c.n_chain_tasks; t = b2_next_task(c, sc)) {
    sc.gsync();
    sc.reset(0);
    int lo = 0, hi = c.b.n_reads;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (c.chain_off[mid] <= t) lo = mid; else hi = mid; }
    const int r = lo;
    const int64_t off = c.seed_off[r];
    const B2Chain& ch = c.chains[off + (t - c.chain_off[r])];
    const int l_seq = c.b.l_seq[r];
    do {
      sc.gsync();
      sc.reset(0);
      uint64_t* srt = (uint64_t*)sc.alloc(sizeof(uint64_t) * (size_t)ch.n, 8);
      B2Reg* av = (B2Reg*)sc.alloc(sizeof(B2Reg) * (size_t)ch.n, 8);
      if (!sc.overflow) {
        int n_av = 0;
        b2_chain2aln(c.opt, c.ix, l_seq, c.b.seq + c.b.seq_off[r], ch, c.cseeds + off + ch.seed_head, av, &n_av, ch.n, srt, sc, 1,
                     c.cand + off + ch.seed_head, c.cand_ok + off + ch.seed_head,
                     c.xext_res ? c.xext_res + 2 * (off + ch.seed_head) : 0);
      }
    } while (b2_retry_big(c, sc));
    b2_end_task(c, sc);
  }
  b2_flush_counters(c, sc);
}

// Serial per-read pass of the extension stage (one lane per read): decides which seeds are extended in the
// reference's order, taking the regions from cand[], then mem_sort_dedup_patch.
B2_KERNEL_LB(128, 4) k_extend(B2Ctx c) {
  B2Scratch sc = b2_lane_scratch(c, B2_GTID());
  for (int r = b2_next_task(c, sc); r < c.b.n_reads; r = b2_next_task(c, sc)) {
    const int64_t off = c.seed_off[r];
    const int nch = c.n_chain[r];
    const int l_seq = c.b.l_seq[r];
    B2Reg* av = c.regs + c.reg_off[r];
    const int cap = (int)(c.reg_off[r + 1] - c.reg_off[r]);
    int n_av = 0;
    if (nch > 0) {
      int max_n = 0;
      for (int i = 0; i < nch; ++i) max_n = max_n > c.chains[off + i].n ? max_n : c.chains[off + i].n;
      do {
        sc.reset(0);
        n_av = 0;
        uint64_t* srt = (uint64_t*)sc.alloc(sizeof(uint64_t) * (size_t)max_n, 8);
        uint8_t* query = (uint8_t*)sc.alloc((size_t)l_seq + 1, 1);
        if (!sc.overflow) {
          const uint8_t* q0 = c.b.seq + c.b.seq_off[r];
          for (int i = 0; i < nch && !sc.overflow; ++i) {
            const B2Chain& ch = c.chains[off + i];
            b2_chain2aln(c.opt, c.ix, l_seq, q0, ch, c.cseeds + off + ch.seed_head, av, &n_av, cap, srt, sc, 2,
                         c.cand + off + ch.seed_head, c.cand_ok + off + ch.seed_head);
          }
          if (!sc.overflow && n_av > 1) {
            for (int i = 0; i < l_seq; ++i) query[i] = q0[i];  // patching may reverse the read in place
            n_av = b2_sort_dedup_patch(c.opt, c.ix, 1, query, n_av, av, sc);
          }
          for (int i = 0; i < n_av; ++i)
            if (av[i].rid >= 0 && c.ix.anns[av[i].rid].is_alt) av[i].is_alt = 1;
        }
      } while (b2_retry_big(c, sc));
      if (b2_end_task(c, sc)) n_av = 0;
    }
    c.n_regs[r] = n_av;
  }
  b2_flush_counters(c, sc);
}

B2_KERNEL k_pestat(B2Ctx c) {
  const int n_pairs = c.b.n_reads >> 1;
  for (int i = B2_GTID(); i < n_pairs; i += B2_GSIZE()) {
    int64_t is = 0;
    int d = b2_pestat_candidate(c.opt, c.ix.l_pac, c.n_regs[2 * i], c.regs + c.reg_off[2 * i], c.n_regs[2 * i + 1],
                                c.regs + c.reg_off[2 * i + 1], &is);
    c.pe_dir[i] = (int8_t)d;
    c.pe_is[i] = is;
  }
}

// The two reads of a pair must carry the same name (after bseq_read's /1 /2 trimming): mem_sam_pe stops the reference's
// run on the first pair that does not (bwa/bwamem_pair.c:355,385) -- the sign of mate files that are out of step.
// *first_bad = lowest such pair of the batch.
B2_KERNEL k_pair_names(B2Ctx c, int* first_bad) {
  const int n_pairs = c.b.n_reads >> 1;
  for (int i = B2_GTID(); i < n_pairs; i += B2_GSIZE()) {
    const int la = c.b.l_name[2 * i];
    bool same = la == c.b.l_name[2 * i + 1];
    const char *a = c.b.names + c.b.name_off[2 * i], *b = c.b.names + c.b.name_off[2 * i + 1];
    for (int k = 0; same && k < la; ++k) same = a[k] == b[k];
    if (!same) B2_ATOMIC_MIN(first_bad, i);
  }
}

// mode 0: count the planned rescue attempts of every pair; mode 1: write them at resc_off[pair]
B2_KERNEL k_resc_plan(B2Ctx c, int mode) {
  const int n_pairs = c.b.n_reads >> 1;
  for (int i = B2_GTID(); i < n_pairs; i += B2_GSIZE()) {
    int n[2], l_seq[2];
    const B2Reg* a[2];
    for (int k = 0; k < 2; ++k) {
      n[k] = c.n_regs[2 * i + k];
      a[k] = c.regs + c.reg_off[2 * i + k];
      l_seq[k] = c.b.l_seq[2 * i + k];
    }
    int cnt = b2_rescue_plan(c.opt, c.ix, c.pes, n, a, l_seq, i, mode ? c.resc_tasks + c.resc_off[i] : 0);
    if (!mode) c.resc_cnt[i] = cnt;
  }
}

// one task group per planned attempt: local SW of the mate against the rescue window
B2_KERNEL k_resc_sw(B2Ctx c) {
  int gid, n_groups;
  B2Scratch sc = b2_group_scratch(c, &gid, &n_groups, B2_G_RESC);
  for (int t = b2_next_task(c, sc); t < c.n_resc; t = b2_next_task(c, sc)) {
    const B2RescTask task = c.resc_tasks[t];
    const int mate = 2 * task.item + (task.end_i ? 0 : 1);
    B2Kswr r;
    do {
      sc.gsync();
      sc.reset(0);
      r = b2_rescue_sw(c.opt, c.ix, task, c.b.seq + c.b.seq_off[mate], sc);
    } while (b2_retry_big(c, sc));
    b2_end_task(c, sc);
    if (sc.lane == 0) c.resc_res[t] = r;
  }
  if (sc.lane == 0 && sc.sw_cells) B2_ATOMIC_ADD(&c.counters[B2_CNT_RESCSW_CELLS], sc.sw_cells);
  b2_flush_counters(c, sc);
}

// ---- speculative global alignments (b2_galn.cuh) ----------------------------------------------------------------
// mode 0: count the alignments worth computing ahead, per band class; mode 1: write them (class-major task ids) and
// enter them into the hash table.  One lane per read; mirrors the entry conditions of mem_reg2aln / bwa_gen_cigar2.
B2_D inline void b2_galn_admit(const B2Ctx& c, int r, const B2Reg* ar, int mode, int* cnt) {
  const int n = c.b.n_reads;
  const int64_t l_pac = c.ix.l_pac;
  if (ar->rb < 0 || ar->re < 0) return;
  const int qb = ar->qb, qe = ar->qe, l_query = qe - qb;
  const int64_t rb = ar->rb, re = ar->re;
  if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac) || re - rb > l_query + 64) return;
  const int w_in = b2_reg2aln_w0(c.opt, ar);
  if (l_query == re - rb && w_in == 0) return;  // gap-free: no DP (bwa/bwa.c:131)
  const int w = b2_cigar_band(c.opt, l_query, (int)(re - rb), w_in);
  const int cls = b2_galn_class(w);
  const int dl = l_query > (int)(re - rb) ? l_query - (int)(re - rb) : (int)(re - rb) - l_query;
  if (cls < 0 || dl > w) return;
  if (mode) {
    const int64_t id = c.galn_base[cls] + c.galn_off[(size_t)cls * (n + 2) + r] + cnt[cls];
    B2GalnTask t;
    t.seq_off = c.b.seq_off[r]; t.rb = rb; t.re = re; t.read = r; t.qb = qb; t.qe = qe; t.w_in = w_in; t.w = w;
    c.galn_tasks[id] = t;
    c.galn_res[id].n_cigar = -1;
    uint32_t h = b2_galn_hash(t.seq_off, qb, qe, rb, re, w_in) & c.galn.mask;
    while (B2_ATOMIC_CAS(&c.galn_table[h], -1, (int32_t)id) != -1) h = (h + 1) & c.galn.mask;
  }
  ++cnt[cls];
}

B2_KERNEL k_galn_plan(B2Ctx c, int mode) {
  const int n = c.b.n_reads;
  const int64_t l_pac = c.ix.l_pac;
  for (int r = B2_GTID(); r < n; r += B2_GSIZE()) {
    int cnt[B2_GALN_CLASSES];
    for (int k = 0; k < B2_GALN_CLASSES; ++k) cnt[k] = 0;
    const B2Reg* a = c.regs + c.reg_off[r];
    const int nr = c.n_regs[r];
    // the two best regions (primary / supplementary candidates), and -- when few enough regions score within
    // XA_drop_ratio of the best for an XA tag to be written at all (mem_gen_alt, bwa/bwamem_extra.c:98-140) -- those too
    int n_plan = nr < B2_GALN_MAXREGS ? nr : B2_GALN_MAXREGS;
    if (nr > B2_GALN_MAXREGS && !(c.opt.flag & B2_F_ALL)) {
      int m = 0;
      while (m < nr && m <= c.opt.max_XA_hits + 1 && (double)a[m].score >= (double)a[0].score * (double)c.opt.XA_drop_ratio) ++m;
      if (m <= c.opt.max_XA_hits + 1 && m > n_plan) n_plan = m;
    }
    for (int k = 0; k < n_plan; ++k) b2_galn_admit(c, r, &a[k], mode, cnt);
    // regions that mate rescue is going to add to this read: the local SW of every planned attempt has run
    // (k_resc_sw), so the region mem_matesw builds from a hit (bwa/bwamem_pair.c:152-172) is known here
    if ((c.opt.flag & B2_F_PE) && c.resc_off && c.n_resc > 0) {
      const int item = r >> 1;
      int planned = 0;
      for (int64_t ti = c.resc_off[item]; ti < c.resc_off[item + 1] && planned < B2_GALN_MAXREGS; ++ti) {
        const B2RescTask t = c.resc_tasks[ti];
        if (2 * t.item + (t.end_i ? 0 : 1) != r) continue;
        const B2Kswr aln = c.resc_res[ti];
        if (!(aln.score >= c.opt.min_seed_len && aln.qb >= 0)) continue;
        const int is_rev = (t.r >> 1 != (t.r & 1));
        const int l_ms = t.l_ms;
        B2Reg b = B2Reg();
        b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
        b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
        b.rb = is_rev ? (l_pac << 1) - (t.rb + aln.te + 1) : t.rb + aln.tb;
        b.re = is_rev ? (l_pac << 1) - (t.rb + aln.tb) : t.rb + aln.te + 1;
        b.score = aln.score;
        b.csub = aln.score2;
        b.secondary = -1;
        b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
        b2_galn_admit(c, r, &b, mode, cnt);
        ++planned;
      }
    }
    if (!mode) for (int k = 0; k < B2_GALN_CLASSES; ++k) c.galn_cnt[(size_t)k * (n + 1) + r] = cnt[k];
  }
}

// one alignment per lane; the lanes of a warp share a pool made of their scratch regions: per-lane staging of the
// two sequences, then the lane-interleaved direction words.
// The four band classes run as ONE launch: blocks [galn_blk[k], galn_blk[k+1]) of the grid serve class k (a class
// alone is a few hundred warps, one or two per SM; side by side they overlap each other's latency).  tid0 / tsize:
// the lane's index and the lane count inside its class's share of the grid.
template <int ND>
B2_D inline void b2_galn_run(const B2Ctx& c, int cls, int tid0, int tsize) {
  const int LW = B2_WARP_LANES;
  const int lane = B2_GTID() & (LW - 1);
  uint8_t* pool = c.scratch + (size_t)(B2_GTID() - lane) * c.scratch_per_lane;
  const int half = (c.max_len + 64 + 15) & ~15;
  uint8_t* q = pool + (size_t)lane * 2 * half;
  uint8_t* tg = q + half;
  uint32_t* zw = (uint32_t*)(pool + (size_t)LW * 2 * half) + lane;
  const int64_t first = c.galn_base[cls], last = first + c.galn_n[cls];
  unsigned long long cells = 0, calls = 0;
  for (int64_t t = first + tid0; t < last; t += tsize) {
    const B2GalnTask task = c.galn_tasks[t];
    B2GalnRes* res = &c.galn_res[t];
    const int qlen = task.qe - task.qb, tlen = (int)(task.re - task.rb);
    if (qlen > c.max_len || tlen > c.max_len + 64) continue;  // (n_cigar stays -1: not available)
    if (b2_get_seq(c.ix.l_pac, c.ix.pac, task.rb, task.re, tg) != tlen) continue;
    const uint8_t* src = c.b.seq + task.seq_off + task.qb;
    if (task.rb >= c.ix.l_pac) {  // reverse both, as bwa_gen_cigar2 does for the reverse strand (bwa/bwa.c:127-130)
      for (int k = 0; k < qlen; ++k) q[k] = src[qlen - 1 - k];
      b2_revseq(tlen, tg);
    } else {
      for (int k = 0; k < qlen; ++k) q[k] = src[k];
    }
    b2_galn_one<ND>(qlen, q, tlen, tg, c.opt.mat, c.opt.o_del, c.opt.e_del, c.opt.o_ins, c.opt.e_ins, task.w, zw, LW, res, &cells);
    ++calls;
  }
  if (calls) {
    B2_ATOMIC_ADD(&c.counters[B2_CNT_GLOB_CELLS], cells); B2_ATOMIC_ADD(&c.counters[B2_CNT_GLOB_CALLS], calls);
    B2_ATOMIC_ADD(&c.counters[B2_CNT_GALN_CELLS], cells);
  }
}
B2_KERNEL k_galn(B2Ctx c) {
#if defined(B2_HOSTSIM)
  const int blk = B2_GTID() / b2_dim.block, bdim = b2_dim.block, tin = B2_GTID() % b2_dim.block;
#else
  const int blk = (int)blockIdx.x, bdim = (int)blockDim.x, tin = (int)threadIdx.x;
#endif
  int cls = 0;
  while (cls < B2_GALN_CLASSES - 1 && blk >= c.galn_blk[cls + 1]) ++cls;
  const int tid0 = (blk - c.galn_blk[cls]) * bdim + tin, tsize = (c.galn_blk[cls + 1] - c.galn_blk[cls]) * bdim;
  if (cls == 0) b2_galn_run<9>(c, 0, tid0, tsize);
  else if (cls == 1) b2_galn_run<17>(c, 1, tid0, tsize);
  else if (cls == 2) b2_galn_run<33>(c, 2, tid0, tsize);
  else b2_galn_run<65>(c, 3, tid0, tsize);
}

B2_KERNEL_LB(128, B2_LB_FINAL) k_final_se(B2Ctx c) {
  int gid, n_groups;
  B2Scratch sc = b2_group_scratch(c, &gid, &n_groups, B2_G_FINAL);
  sc.galn = &c.galn;
  for (int r = b2_next_task(c, sc); r < c.b.n_reads; r = b2_next_task(c, sc)) {
    int err = 0;
    const int n = c.n_regs[r];
    B2Out out;
    do {
      sc.gsync();
      sc.reset(0);
      err = 0;
      B2Reg* a = (B2Reg*)sc.alloc(sizeof(B2Reg) * (size_t)(n > 0 ? n : 1), 8);
      int* z = (int*)sc.alloc(sizeof(int) * (size_t)(n + 1), 4);
      out.p = c.slots + (size_t)r * c.slot_size; out.cap = c.slot_size; out.len = 0;
      out.lane = sc.lane; out.gsize = sc.gsize;
      if (!sc.overflow) {
        for (int i = 0; i < n; ++i) a[i] = c.regs[c.reg_off[r] + i];
        B2Read s = b2_make_read(c.b, r, 1);
        b2_mark_primary_se(c.opt, n, a, c.n_processed + r, z, sc);
        b2_reg2sam(c.opt, c.ix, c.tb, c.fmt, out, s, n, a, 0, 0, sc, &err);
      }
    } while (b2_retry_big(c, sc));
    b2_end_task(c, sc);
    if (err) B2_ATOMIC_OR(c.flags, B2_FLAG_ERR_TABLE);
    if (out.len > out.cap) B2_ATOMIC_OR(c.flags, B2_FLAG_OVF_SLOT);
    c.sam_len[r] = (int32_t)out.len;
  }
  b2_flush_counters(c, sc);
}

// Pairs are finalised in two launches (c.fin_mode).  The few pairs with many candidate regions -- reads in repeats:
// dozens of rescue attempts, region lists re-sorted after each, XA alignments -- take 10-20x the time of an ordinary
// pair; inside a warp of sub-warp groups the other groups wait for such a pair at the end of every trip of the task
// loop.  k_final_plan lists them (fin_list, flags in fin_heavy); k_final_pe_heavy gives each a whole warp
// (B2_G_HEAVY lanes for its DP, sorts' data movement and text), k_final_pe then runs the ordinary pairs in
// B2_G_FINAL-lane groups, skipping the flagged ones.  fin_mode 0: one launch over all pairs (single-end batches,
// splitting disabled).
// 1 iff mem_reg2aln of this region is going to run a global alignment inside the finalisation kernel: the region is not
// gap-free and either no alignment was computed ahead for it (band wider than k_galn's classes, lengths too different) or
// the one that was scores below truesc - a, so that the band is doubled and the DP repeated (bwa/bwamem.c:1072-1086)
B2_D inline int b2_final_needs_dp(const B2Ctx& c, int r, const B2Reg* ar) {
  if (ar->rb < 0 || ar->re < 0) return 0;
  const int qb = ar->qb, qe = ar->qe;
  const int64_t rb = ar->rb, re = ar->re;
  if (qe - qb <= 0 || rb >= re || (rb < c.ix.l_pac && re > c.ix.l_pac)) return 0;
  const int w2 = b2_reg2aln_w0(c.opt, ar);
  if (qe - qb == re - rb && w2 == 0) return 0;
  const B2GalnRes* pre = b2_galn_lookup(&c.galn, c.b.seq + c.b.seq_off[r], qb, qe, rb, re, w2);
  if (pre == 0) return 1;
  return w2 != c.opt.w << 2 && pre->score < ar->truesc - c.opt.a;
}

B2_KERNEL k_final_plan(B2Ctx c) {
  const int n_pairs = c.b.n_reads >> 1;
  for (int i = B2_GTID(); i < n_pairs; i += B2_GSIZE()) {
    int heavy = c.n_regs[2 * i] + c.n_regs[2 * i + 1] >= c.fin_heavy_min;
    // a pair whose leading regions need a DP in the kernel also gets a warp: 32 lanes for the alignment, and the other
    // pairs of a 4-lane-group warp do not wait for it
    for (int k = 0; k < 2 && !heavy; ++k) {
      const int r = 2 * i + k, n = c.n_regs[r] < 2 ? c.n_regs[r] : 2;
      for (int j = 0; j < n && !heavy; ++j) heavy = b2_final_needs_dp(c, r, c.regs + c.reg_off[r] + j);
    }
    c.fin_heavy[i] = (uint8_t)heavy;
    if (heavy) c.fin_list[B2_ATOMIC_ADD(c.work + 3, 1)] = i;
  }
}

template <int G>
B2_D inline void b2_final_pe_body(const B2Ctx& c) {
  int gid, n_groups;
  B2Scratch sc = b2_group_scratch(c, &gid, &n_groups, G);
  sc.galn = &c.galn;
#if defined(__CUDA_ARCH__)
  extern __shared__ uint4 b2_fin_smem[];   // c.fin_fast bytes per task group (B2Scratch::alloc_fast)
  sc.fast = (uint8_t*)b2_fin_smem + (size_t)(threadIdx.x / G) * (size_t)c.fin_fast;
  sc.fast_cap = (unsigned)c.fin_fast;
#endif
  const int n_pairs = c.b.n_reads >> 1;
  const int n_tasks = c.fin_mode == 2 ? c.work[3] : n_pairs;
  sc.prof = c.dbg ? c.dbg + 4 * (size_t)n_pairs : 0;
  for (int t = b2_next_task(c, sc); t < n_tasks; t = b2_next_task(c, sc)) {
    int i = t;
    if (c.fin_mode == 2) i = c.fin_list[t];
    else if (c.fin_mode == 1 && c.fin_heavy[t]) continue;
#if defined(__CUDA_ARCH__)
    const long long t0 = clock64();
    const unsigned sw0 = sc.sw_calls, gl0 = sc.glob_calls;
#endif
    int err = 0;
    int n[2], cap[2];
    B2Out out;
    do {   // (a second time in a slot of the outlier pool if the group's own region is too small for this pair)
      sc.gsync();
      sc.reset(0);
      sc.tick(-1);
      err = 0;
      B2Reg* a[2];
      B2Read s[2];
      n[0] = c.n_regs[2 * i]; n[1] = c.n_regs[2 * i + 1];
      for (int k = 0; k < 2; ++k) {
        int other = n[!k] < c.opt.max_matesw ? n[!k] : c.opt.max_matesw;
        cap[k] = n[k] + ((c.opt.flag & B2_F_NO_RESCUE) ? 0 : 4 * other) + 1;
        a[k] = (B2Reg*)sc.alloc(sizeof(B2Reg) * (size_t)cap[k], 8);
        s[k] = b2_make_read(c.b, 2 * i + k, 1);
      }
      out.p = c.slots + (size_t)i * c.slot_size; out.cap = c.slot_size; out.len = 0;
      out.lane = sc.lane; out.gsize = sc.gsize;
      if (!sc.overflow) {
        for (int k = 0; k < 2; ++k)
          for (int j = 0; j < n[k]; ++j) a[k][j] = c.regs[c.reg_off[2 * i + k] + j];
        B2RescCache cache;
        cache.n = 0; cache.tasks = 0; cache.res = 0;
        if (c.resc_off) {
          cache.n = (int)(c.resc_off[i + 1] - c.resc_off[i]);
          cache.tasks = c.resc_tasks + c.resc_off[i];
          cache.res = c.resc_res + c.resc_off[i];
        }
        sc.tick(7);
        b2_sam_pe(c.opt, c.ix, c.tb, c.fmt, c.pes, (uint64_t)((c.n_processed >> 1) + i), s, a, n, cap, out, sc, &err, &cache);
      }
    } while (b2_retry_big(c, sc));
    b2_end_task(c, sc);
    if (err) B2_ATOMIC_OR(c.flags, B2_FLAG_ERR_TABLE);
    if (out.len > out.cap) B2_ATOMIC_OR(c.flags, B2_FLAG_OVF_SLOT);
    c.sam_len[i] = (int32_t)out.len;
#if defined(__CUDA_ARCH__)
    if (c.dbg && sc.lane == 0) {
      c.dbg[4 * i] = clock64() - t0; c.dbg[4 * i + 1] = sc.sw_calls - sw0; c.dbg[4 * i + 2] = sc.glob_calls - gl0;
      c.dbg[4 * i + 3] = (long long)c.n_regs[2 * i] << 32 | (unsigned)c.n_regs[2 * i + 1];
    }
#endif
  }
  b2_flush_counters(c, sc);
}

B2_KERNEL_LB(128, B2_LB_FINAL) k_final_pe(B2Ctx c) { b2_final_pe_body<B2_G_FINAL>(c); }
B2_KERNEL_LB(128, B2_LB_FINAL) k_final_pe_heavy(B2Ctx c) { b2_final_pe_body<B2_G_HEAVY>(c); }

// exclusive scan int32 -> int64, out[n] = total.  One block.
B2_KERNEL k_scan(const int32_t* in, int64_t* out, int n) {
#if defined(B2_HOSTSIM)
  if (B2_GTID() != 0) return;
  int64_t s = 0;
  for (int i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
  out[n] = s;
#else
  // one block; 4 consecutive elements per thread and trip (4096 per trip, coalesced), thread-local prefix, then warp
  // shuffles over the per-thread sums and one pass over the warp totals; the running total is carried on
  __shared__ int64_t wsum[32];
  __shared__ int64_t chunk_total;
  const int t = threadIdx.x, T = blockDim.x, lane = t & 31, warp = t >> 5, n_warps = T >> 5;
  int64_t base = 0;
  for (int c0 = 0; c0 < n; c0 += 4 * T) {
    const int i = c0 + 4 * t;
    int64_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i + k < n ? (int64_t)in[i + k] : 0;
    const int64_t mine = v[0] + v[1] + v[2] + v[3];
    int64_t x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const int64_t wv = lane < n_warps ? wsum[lane] : 0;
      int64_t w = wv;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      wsum[lane] = w - wv;                 // exclusive prefix of the warp totals
      if (lane == 31) chunk_total = w;
    }
    __syncthreads();
    int64_t run = base + wsum[warp] + x - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i + k < n) out[i + k] = run; run += v[k]; }
    base += chunk_total;
    __syncthreads();
  }
  if (t == 0) out[n] = base;
#endif
}

// ordered compaction of SAM slots: 16-byte pieces, one per lane
B2_KERNEL k_gather(B2Ctx c, int n_items) {
  const int pieces = (c.slot_size + 15) / 16;
  const int64_t total = (int64_t)n_items * pieces;
  for (int64_t t = B2_GTID(); t < total; t += B2_GSIZE()) {
    const int item = (int)(t / pieces), pc = (int)(t % pieces);
    const int len = c.sam_len[item];
    const int b = pc * 16;
    if (b >= len) continue;
    const int e = b + 16 < len ? b + 16 : len;
    const char* src = c.slots + (size_t)item * c.slot_size;
    char* dst = c.sam_out + c.sam_off[item];
    for (int k = b; k < e; ++k) dst[k] = src[k];
  }
}

// ---------------------------------------------------------------------------------------------
// Read ingest on the device (bseq_read's per-record work, bwa/bwa.c:42-75): the host ships the raw FASTQ bytes of a
// batch plus one descriptor per record; these kernels trim the read-number suffix of the name (trim_readno,
// bwa/bwa.c:27-31), split off the comment (kseq: the header up to the first white space is the name, bwa/kseq.h:181-193),
// translate the bases with nst_nt4_table (bwa/bwamem.c:1025) and lay the batch out as B2Batch.
struct B2Ingest {
  const uint8_t* raw;        // staged file bytes
  const B2Rec* recs[2];      // per input file
  int64_t adj[2];            // raw + rec.off + adj[f] = address of the record's '@'
  int n_files, n_reads, copy_comment;
  uint8_t* seq; char* qual; uint8_t* has_qual; char* names; char* comments;
  const int64_t *seq_off, *name_off, *com_off;
  int32_t *l_seq, *l_name, *l_com, *com_cnt;
};

B2_HD inline const uint8_t* b2_ingest_rec(const B2Ingest& g, int r, B2Rec* rec) {
  const int f = g.n_files == 2 ? (r & 1) : 0;
  const int i = g.n_files == 2 ? (r >> 1) : r;
  *rec = g.recs[f][i];
  return g.raw + (int64_t)rec->off + g.adj[f];
}

B2_KERNEL k_ingest_len(B2Ingest g) {
  for (int r = B2_GTID(); r < g.n_reads; r += B2_GSIZE()) {
    B2Rec rec;
    const uint8_t* p = b2_ingest_rec(g, r, &rec);
    uint32_t l = rec.l_name;
    if (l > 2 && p[1 + l - 2] == '/' && p[1 + l - 1] >= '0' && p[1 + l - 1] <= '9') l -= 2;
    g.l_seq[r] = (int32_t)rec.l_seq;
    g.l_name[r] = (int32_t)l;
    const int lc = (g.copy_comment && rec.l_hdr > rec.l_name + 1) ? (int)(rec.l_hdr - rec.l_name - 1) : -1;
    g.l_com[r] = lc;
    g.com_cnt[r] = lc > 0 ? lc : 0;
    g.has_qual[r] = 1;
  }
}

B2_KERNEL k_ingest_copy(B2Ingest g) {
  const int64_t total = (int64_t)g.n_reads * 4;
  for (int64_t t = B2_GTID(); t < total; t += B2_GSIZE()) {
    const int r = (int)(t >> 2), part = (int)(t & 3);
    B2Rec rec;
    const uint8_t* p = b2_ingest_rec(g, r, &rec);
    if (part == 0) {
      const uint8_t* src = p + 1 + rec.l_hdr + 1;
      uint8_t* dst = g.seq + g.seq_off[r];
      for (uint32_t k = 0; k < rec.l_seq; ++k) {
        const uint32_t ch = src[k] | 0x20u;  // nst_nt4_table: A/a 0, C/c 1, G/g 2, T/t 3, everything else 4
        dst[k] = ch == 'a' ? 0 : ch == 'c' ? 1 : ch == 'g' ? 2 : ch == 't' ? 3 : 4;
      }
    } else if (part == 1) {
      const uint8_t* src = p + rec.qual_rel;
      char* dst = g.qual + g.seq_off[r];
      for (uint32_t k = 0; k < rec.l_seq; ++k) dst[k] = (char)src[k];
    } else if (part == 2) {
      const uint8_t* src = p + 1;
      char* dst = g.names + g.name_off[r];
      const int l = g.l_name[r];
      for (int k = 0; k < l; ++k) dst[k] = (char)src[k];
    } else {
      const int l = g.l_com[r];
      if (l > 0) {
        const uint8_t* src = p + 1 + rec.l_name + 1;
        char* dst = g.comments + g.com_off[r];
        for (int k = 0; k < l; ++k) dst[k] = (char)src[k];
      }
    }
  }
}
