// Host-side interface of the GPU batch engine (C++; the C-ABI in include/b200bwa.h sits on top).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "b2_types.h"

// Host copy of one index (files <prefix>.bwt/.sa/.pac/.ann/.amb[/.alt], bwa/bwa.c:252-279).
struct B2IndexHost {
  std::vector<uint32_t> bwt;   // payload after the 40-byte header
  std::vector<uint64_t> sa;    // sa[0] = -1
  std::vector<uint8_t> pac;
  std::vector<B2Ann> anns;
  std::string names;           // name/anno blob
  std::vector<std::string> ann_names;
  uint64_t primary, L2[5], seq_len;
  size_t bwt_words = 0, n_sa = 0;  // sizes of the device images (valid even when the vectors are released)
  int sa_intv;
  int64_t l_pac;
  int n_seqs, n_holes;
};

// One -K batch as cut by the reader (bseq_read, bwa/bwa.c:42-75).
struct B2HostBatch {
  int n_reads = 0, max_len = 0;
  int64_t n_bases = 0;
  int64_t n_processed = 0;       // absolute index of the first read (hash tie-breaks depend on it)
  std::vector<uint8_t> seq;      // base codes 0..4
  std::vector<char> qual;
  std::vector<uint8_t> has_qual;
  std::vector<char> names, comments;
  std::vector<int64_t> seq_off, name_off, com_off;
  std::vector<int32_t> l_seq, l_name, l_com;
  void clear() {
    n_reads = max_len = 0; n_bases = 0;
    seq.clear(); qual.clear(); has_qual.clear(); names.clear(); comments.clear();
    seq_off.clear(); name_off.clear(); com_off.clear(); l_seq.clear(); l_name.clear(); l_com.clear();
  }
};

// One -K batch as raw file bytes: per input file the records' descriptors and the mapping they point into.  The engine
// copies the byte range the records span to the device and parses / trims / encodes there (k_ingest_*).
// Paired-end input from two files: read 2i = record i of file 0, read 2i+1 = record i of file 1.
struct B2RawInput {
  int n_files = 0, n_reads = 0, max_len = 0, copy_comment = 0;
  int64_t n_bases = 0, n_processed = 0;
  const unsigned char* base[2] = {0, 0};
  const B2Rec* recs[2] = {0, 0};
  size_t n_recs[2] = {0, 0};
};

struct B2EngineConfig {
  int device = 0;
  int lane_grid = 148 * 4, lane_block = 128;   // lane-per-read kernels (persistent, grid-stride)
  int seed_grid = 148 * 16, seed_block = 128;  // seeding: one task per lane; grid = min(tasks, SMs x blocks per SM, seed_grid)
  int seed_blocks_per_sm = 4;                  // shared memory is split for this many resident seeding blocks per SM
  int seed3_blocks_per_sm = 8;                 // forward-only pass: no stack, fewer registers
  int final_fast_bytes = 512;                  // shared memory per task group of k_final_pe: read copy + reference window of an alignment
  int final_heavy_regs = 4;                    // pairs with at least this many regions (both mates) are finalised a warp each, in their own launch (0: no split)
  int seed_smem_entries = 0;                   // interval-stack entries per lane in shared memory (0: what fits)
  size_t scratch_per_lane = 16 << 10;          // every lane's (task group's) own scratch region: sized for the ordinary read / pair
  size_t big_pool_bytes = (size_t)1 << 30;     // outlier pool per stream set: tasks that overflow their region run again in one of its slots (0: none)
  size_t big_slot_bytes = 256 << 10;           // initial slot size of the pool
  int no_xext = 0;                             // 1: skip the inter-task seed extensions (B200BWA_NO_XEXT)
  int no_ext_chain = 0;                        // 1: skip the per-chain extension kernel, the per-read pass computes in place (B200BWA_NO_EXT_CHAIN)
  int no_resc_plan = 0;                        // 1: no planned mate-rescue SW tasks, the finalisation kernel runs them itself (B200BWA_NO_RESC_PLAN)
  int no_sa_plan = 0;                          // 1: k_sa_search instead of k_sa_plan + k_sa (B200BWA_NO_SA_PLAN)
  int no_galn = 0;                             // 1: skip the speculative inter-task global alignments (B200BWA_NO_GALN)
  size_t max_scratch_bytes = (size_t)16 << 30;    // per stream/buffer set; a stage that needs more per lane runs with fewer lanes
  int n_workers = 6;                           // independent stream/buffer sets (batches in flight)
  int cap_intv = 64;
  int slot_per_read = 0;                       // 0: derived from the longest read of the batch
  int verbose = 3;
};

struct B2StageTimes {  // milliseconds, device time (CUDA events) of the last batch
  float h2d = 0, seed = 0, sa = 0, chain = 0, extend = 0, pestat = 0, rescue = 0, final_ = 0, gather = 0, d2h = 0, total = 0;
  float k_xext = 0, k_resc_sw = 0, k_galn = 0;   // the DP kernels proper (last launch of each in the batch)
  float stage_host = 0;   // raw ingest: host copy of the file bytes into the pinned staging buffer (wall ms)
  int launches = 0;
  int64_t n_seed_tasks = 0, n_regs = 0, sam_bytes = 0, n_resc = 0;
};

class B2EngineImpl;

class B2Engine {
 public:
  B2Engine();
  ~B2Engine();
  // uploads the index; returns 0 on success
  int init(const B2IndexHost& idx, const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg);
  // adopt an index image that already lives in device memory (NCCL-broadcast by the caller)
  int init_device_index(const B2IndexHost& meta, const void* d_bwt, const void* d_sa, const void* d_pac,
                        const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg);
  // replicate the index image of `src` (an engine of another device of this process) into this engine's device:
  // device-to-device copies over NVLink (cudaMemcpyPeerAsync), no second trip through host memory
  int init_from_peer(const B2IndexHost& meta, B2Engine& src, const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg);
  int device() const;
  // Aligns one batch and appends its SAM text to `sam`.  PE iff opt.flag & B2_F_PE.  pes0: user-supplied
  // insert-size distribution (-I) or null.  `worker` selects one of the independent stream/buffer sets.
  int process(const B2HostBatch& b, const B2Pes* pes0, std::string& sam, int worker = 0);
  // the same two calls for a batch given as raw FASTQ bytes (packing on the device)
  int process_raw(const B2RawInput& in, const B2Pes* pes0, std::string& sam, int worker);
  int process_raw_view(const B2RawInput& in, const B2Pes* pes0, int worker, const char** data, size_t* len);
  // same, leaving the SAM text in one of the worker's two pinned host buffers (used in turn): *data stays valid until
  // the SECOND next call on `worker`, so the caller can write batch i out while batch i+1 is computed
  int next_view_slot(int worker) const;   // which of the two buffers the next *_view call on `worker` will fill
  int process_view(const B2HostBatch& b, const B2Pes* pes0, int worker, const char** data, size_t* len);
  // one batch forced single-end (paired = 0) or paired-end (1), also returning the SAM bytes of every read / pair:
  // the two halves of a smart-pairing batch (bwa/fastmap.c:64-83)
  int process_items(const B2HostBatch& b, const B2Pes* pes0, int worker, int paired, std::string& sam,
                    std::vector<int32_t>& item_len);
  // device-resident variant used by the benchmark: upload once, run the kernels many times
  int upload(const B2HostBatch& b, int worker = 0);
  int run_resident(const B2Pes* pes0, std::string* sam, int worker = 0);
  // reads [first, first+count) of the batch resident on src_worker, run on `worker`'s stream/buffers
  // paired: 0 / 1 = this range single-end / paired-end, -1 = as the options say
  int run_resident_range(const B2Pes* pes0, std::string* sam, int worker, int src_worker, int first, int count, int paired = -1);
  // per-call options (a cached index serves successive calls with different bwa-mem options)
  void set_options(const B2Opt& opt, const char* rg_id, int verbose);
  const B2Opt& options() const;
  const B2StageTimes& times(int worker = 0) const;
  // cumulative algorithmic-work counters of a worker (see B2_CNT_* in b2_kernels.cuh) plus copied bytes
  int counters(int worker, unsigned long long out[16], unsigned long long* h2d_bytes, unsigned long long* d2h_bytes, int reset);
  const char* last_error() const;
  int n_workers() const;
  // raw device pointers of the index image (for broadcasting it to peer GPUs)
  void index_device_ptrs(void** bwt, size_t* bwt_bytes, void** sa, size_t* sa_bytes, void** pac, size_t* pac_bytes);
  // stage dumps for the parity tests (valid after process/run_resident on that worker)
  int dump_intervals(int worker, std::vector<int32_t>& n_intv, std::vector<B2Intv>& intv);
  int dump_regs(int worker, std::vector<int32_t>& n_regs, std::vector<int64_t>& reg_off, std::vector<B2Reg>& regs);

 private:
  B2EngineImpl* impl_;
};

// number of usable CUDA devices (0 when there is none)
int b2_device_count();

// mem_pestat second half (bwa/bwamem_pair.c:66-108) on the host: candidates -> pes[4]
void b2_pestat_host(const B2Opt& opt, int n_pairs, const int8_t* dir, const int64_t* is, B2Pes pes[4], int verbose);
