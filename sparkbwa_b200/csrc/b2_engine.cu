// GPU batch engine: owns the device-resident index image and per-worker stream/buffer sets, and
// stages one -K batch through the kernels of b2_kernels.cuh (the unit the reference computes in
// mem_process_seqs, bwa/bwamem.c:1172-1201).  Capacity overflows reported by a stage are answered
// by re-running that (idempotent) stage with larger buffers; anything else fails loudly.
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>
#include "b2_engine.h"
#include "b2_kernels.cuh"

#if defined(B2_HOSTSIM)
thread_local B2Dim b2_dim;
#if defined(B2_GALN_STATS)
long b2_galn_stats[16];
struct B2GalnStatsPrinter { ~B2GalnStatsPrinter() { fprintf(stderr, "[D::galn] reg2aln attempts: no-DP %ld, hit %ld, retry-miss %ld, wide-miss %ld, other-miss %ld; missed cells %ld\n",
  b2_galn_stats[0], b2_galn_stats[1], b2_galn_stats[2], b2_galn_stats[3], b2_galn_stats[4], b2_galn_stats[8]); } } b2_galn_stats_printer;
#endif
#endif

#include <atomic>
#include <chrono>
#include <mutex>
std::atomic<unsigned long long> b2_g_h2d_bytes{0}, b2_g_d2h_bytes{0}, b2_g_launches{0};

namespace {

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n, bool exact = false) {
    if (n <= cap) return 0;
    // generous headroom: a cudaFree/cudaMalloc pair synchronises the whole device, i.e. stalls every other stream,
    // and per-batch sizes (seeds, regions, rescue windows, SAM bytes) keep creeping up for the first dozen batches
    size_t want = exact ? n : n + n / 2 + 4096;
    b2_dev_free(p);
    p = nullptr; cap = 0;
    void* q = nullptr;
    if (b2_dev_malloc(&q, want * sizeof(T))) return 1;
    p = (T*)q; cap = want;
    return 0;
  }
  void release() { b2_dev_free(p); p = nullptr; cap = 0; }
};

struct Timer {
#if !defined(B2_HOSTSIM)
  cudaEvent_t e[32];
  int n = 0;
  cudaStream_t s;
  void init(cudaStream_t st) { s = st; for (auto& x : e) cudaEventCreate(&x); }
  void destroy() { for (auto& x : e) cudaEventDestroy(x); }
  void reset() { n = 0; }
  int mark() { cudaEventRecord(e[n], s); return n++; }
  float ms(int a, int b) { float t = 0; cudaEventElapsedTime(&t, e[a], e[b]); return t; }
#else
  void init(int) {}
  void destroy() {}
  void reset() {}
  int mark() { return 0; }
  float ms(int, int) { return 0.f; }
#endif
};

struct Worker {
  b2_stream_t stream{};
  Timer timer;
  // batch
  DBuf<uint8_t> seq; DBuf<char> qual; DBuf<uint8_t> has_qual; DBuf<char> names, comments;
  DBuf<int64_t> seq_off, name_off, com_off; DBuf<int32_t> l_seq, l_name, l_com;
  // stages
  DBuf<B2Intv> intv, seed_scratch; DBuf<uint32_t> seed_spill; DBuf<int32_t> n_intv, n_tasks; DBuf<int64_t> seed_off;
  DBuf<B2Seed> raw, cseeds; DBuf<int32_t> raw_rid; DBuf<B2Chain> chains; DBuf<int32_t> n_chain, reg_cap;
  DBuf<int64_t> reg_off; DBuf<B2Reg> regs; DBuf<int32_t> n_regs;
  DBuf<uint8_t> scratch; DBuf<int8_t> pe_dir; DBuf<int64_t> pe_is; DBuf<double> pairtab;
  DBuf<char> slots, sam_out; DBuf<int32_t> sam_len; DBuf<int64_t> sam_off;
  DBuf<int> flags;
  DBuf<uint8_t> fin_heavy; DBuf<int32_t> fin_list;
  DBuf<int32_t> sa_task_r; DBuf<uint32_t> sa_task_iv; DBuf<int32_t> dbg_seed;
  DBuf<long long> dbg;
  DBuf<int64_t> chain_off; DBuf<B2Reg> cand; DBuf<uint8_t> cand_ok;
  DBuf<int32_t> resc_cnt; DBuf<int64_t> resc_off; DBuf<B2RescTask> resc_tasks; DBuf<B2Kswr> resc_res;
  DBuf<B2XextTask> xext_tasks; DBuf<B2XextRes> xext_res; DBuf<int32_t> xext_key, xext_hist, xext_perm; DBuf<int64_t> xext_hoff;
  DBuf<int32_t> galn_cnt, galn_table; DBuf<int64_t> galn_off; DBuf<B2GalnTask> galn_tasks; DBuf<B2GalnRes> galn_res;
  DBuf<unsigned long long> counters;
  unsigned long long h_counters[16] = {0};
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;
  size_t scratch_per_lane = 0, big_slot = 0;
  DBuf<uint8_t> big;   // outlier pool
  int clean_batches = 0, decay_wait = 4;   // batches without a scratch overflow / how many of them before the size is halved again
  bool tried_decay = false;
  int cap_intv = 0, slot_size = 0;
  // resident batch metadata
  int n_reads = 0, max_len = 0;
  int64_t n_processed = 0;
  bool resident = false;
  int n_reads_run = 0;
  B2StageTimes times;
  std::vector<int8_t> h_dir; std::vector<int64_t> h_is; std::vector<char> h_sam;
  // two pinned SAM buffers used in turn: the caller may still be writing batch i's text to the output file while
  // batch i+1 is computed and downloaded (process_view's pointer stays valid until the second next call)
  void* pin_sams[2] = {nullptr, nullptr}; size_t pin_sam_caps[2] = {0, 0}; int sam_flip = 0;
  void* pin_sam = nullptr; size_t last_sam_bytes = 0;
  // raw-FASTQ ingest: pinned staging of the batch's file bytes + record descriptors, and its device image
  void* pin_in = nullptr; size_t pin_in_cap = 0;
  DBuf<uint8_t> raw_in; DBuf<int32_t> com_cnt;
  int ev_in0 = -1, ev_in1 = -1;
  float stage_host_ms = 0;
};

}  // namespace

class B2EngineImpl {
 public:
  B2EngineConfig cfg;
  B2Opt opt;
  B2Index ix{};  // device pointers
  DBuf<uint32_t> d_bwt; DBuf<uint64_t> d_sa; DBuf<uint8_t> d_pac; DBuf<B2Ann> d_anns; DBuf<char> d_names;
  DBuf<double> d_logtab; DBuf<char> d_rg;
  bool owns_index = true;
  size_t bwt_bytes = 0, sa_bytes = 0, pac_bytes = 0;
  int n_log = 1 << 21;   // log(i) for every length a read of up to ~1 Mbp can produce (region spans reach twice the read length)
  B2Fmt fmt{};
  std::vector<Worker> workers;
  std::atomic<size_t> grown_big_slot{0};
  std::atomic<size_t> grown_scratch{0};            // sizes some worker had to grow to, taken over by the others (run())
  std::atomic<int> grown_cap_intv{0}, grown_slot{0};
  std::string err;
  mutable std::mutex err_mu;  // fail() is called from every GPU worker thread

  int fail(const std::string& m) {
    { std::lock_guard<std::mutex> lk(err_mu); err = m; }
    if (cfg.verbose >= 1) fprintf(stderr, "[E::b200bwa] %s\n", m.c_str());
    return 1;
  }
  std::string last_error() const { std::lock_guard<std::mutex> lk(err_mu); return err; }

  int common_init(const B2IndexHost& h, const B2Opt& o, const char* rg_id, const B2EngineConfig& c) {
    cfg = c; opt = o;
#if !defined(B2_HOSTSIM)
    if (cudaSetDevice(cfg.device) != cudaSuccess) return fail("cudaSetDevice failed (no usable CUDA device)");
    // a per-device function attribute, the same value from every engine of the process
    if (cudaFuncSetAttribute(k_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail("cudaFuncSetAttribute(k_seed) failed");
    if (cudaFuncSetAttribute(k_xext_run, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail("cudaFuncSetAttribute(k_xext_run) failed");
#endif
    if (h.seq_len >= (1ull << 36)) return fail("index too large: the seeding kernel packs interval bounds into 36 bits");
    ix.primary = h.primary; ix.seq_len = h.seq_len;
    for (int i = 0; i < 5; ++i) ix.L2[i] = h.L2[i];
    ix.bwt_size = h.bwt_words ? h.bwt_words : h.bwt.size(); ix.n_sa = h.n_sa ? h.n_sa : h.sa.size();
    ix.sa_intv = h.sa_intv;
    ix.sa_shift = 0; while ((1 << ix.sa_shift) < h.sa_intv) ++ix.sa_shift;
    ix.l_pac = h.l_pac; ix.n_seqs = h.n_seqs;
    ix.max_name_len = 0;
    for (auto& a : h.anns) ix.max_name_len = std::max(ix.max_name_len, a.name_len);
    if (d_anns.ensure(h.anns.size()) || d_names.ensure(h.names.size() + 1)) return fail("device allocation failed (contig table)");
    b2_h2d(d_anns.p, h.anns.data(), h.anns.size() * sizeof(B2Ann), 0);
    b2_h2d(d_names.p, h.names.data(), h.names.size(), 0);
    ix.anns = d_anns.p; ix.names = d_names.p;
    std::vector<double> lt(n_log);
    lt[0] = log(0.0);
    for (int i = 1; i < n_log; ++i) lt[i] = log((double)i);
    if (d_logtab.ensure(n_log)) return fail("device allocation failed (log table)");
    b2_h2d(d_logtab.p, lt.data(), n_log * sizeof(double), 0);
    if (set_rg(rg_id)) return 1;
    b2_sync(0);
    workers.resize(cfg.n_workers < 1 ? 1 : (cfg.n_workers > 8 ? 8 : cfg.n_workers));
    for (auto& w : workers) {
#if !defined(B2_HOSTSIM)
      if (cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
#endif
      w.timer.init(w.stream);
      w.scratch_per_lane = cfg.scratch_per_lane;
      w.big_slot = std::min(cfg.big_slot_bytes, std::min(cfg.big_pool_bytes ? cfg.big_pool_bytes : cfg.big_slot_bytes, cfg.max_scratch_bytes));
      w.cap_intv = cfg.cap_intv;
      if (w.flags.ensure(8) || w.counters.ensure(16)) return fail("device allocation failed");
      b2_dev_memset(w.counters.p, 0, 16 * sizeof(unsigned long long), w.stream);
    }
    return 0;
  }

  int set_rg(const char* rg_id) {
    size_t l_rg = rg_id ? strlen(rg_id) : 0;
    if (d_rg.ensure(l_rg + 1)) return fail("device allocation failed");
    if (l_rg) b2_h2d(d_rg.p, rg_id, l_rg, 0);
    b2_sync(0);
    fmt.rg_id = d_rg.p; fmt.l_rg = (int)l_rg;
    return 0;
  }

  int init(const B2IndexHost& h, const B2Opt& o, const char* rg_id, const B2EngineConfig& c) {
    cfg = c;
#if !defined(B2_HOSTSIM)
    if (cudaSetDevice(c.device) != cudaSuccess) return fail("cudaSetDevice failed (no usable CUDA device)");
#endif
    bwt_bytes = h.bwt.size() * 4; sa_bytes = h.sa.size() * 8; pac_bytes = h.pac.size();
    if (d_bwt.ensure(h.bwt.size() + 16) || d_sa.ensure(h.sa.size()) || d_pac.ensure(h.pac.size() + 8))
      return fail("device allocation failed (index image)");
    b2_h2d(d_bwt.p, h.bwt.data(), bwt_bytes, 0);
    b2_h2d(d_sa.p, h.sa.data(), sa_bytes, 0);
    b2_h2d(d_pac.p, h.pac.data(), pac_bytes, 0);
    ix.bwt = d_bwt.p; ix.sa = d_sa.p; ix.pac = d_pac.p;
    return common_init(h, o, rg_id, c);
  }

  // index image copied device to device from an engine of another GPU of this process
  int init_peer(const B2IndexHost& h, B2EngineImpl& src, const B2Opt& o, const char* rg_id, const B2EngineConfig& c) {
    cfg = c;
    bwt_bytes = src.bwt_bytes; sa_bytes = src.sa_bytes; pac_bytes = src.pac_bytes;
#if !defined(B2_HOSTSIM)
    if (cudaSetDevice(c.device) != cudaSuccess) return fail("cudaSetDevice failed (no usable CUDA device)");
    if (c.device != src.cfg.device) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, c.device, src.cfg.device) == cudaSuccess && can)
        if (cudaDeviceEnablePeerAccess(src.cfg.device, 0) != cudaSuccess) cudaGetLastError();  // already enabled: fine
    }
#endif
    if (d_bwt.ensure(bwt_bytes / 4 + 16) || d_sa.ensure(sa_bytes / 8 + 1) || d_pac.ensure(pac_bytes + 8))
      return fail("device allocation failed (index image)");
#if !defined(B2_HOSTSIM)
    if (cudaMemcpyPeerAsync(d_bwt.p, c.device, src.ix.bwt, src.cfg.device, bwt_bytes, 0) != cudaSuccess ||
        cudaMemcpyPeerAsync(d_sa.p, c.device, src.ix.sa, src.cfg.device, sa_bytes, 0) != cudaSuccess ||
        cudaMemcpyPeerAsync(d_pac.p, c.device, src.ix.pac, src.cfg.device, pac_bytes, 0) != cudaSuccess)
      return fail("peer copy of the index image failed");
#else
    memcpy(d_bwt.p, src.ix.bwt, bwt_bytes); memcpy(d_sa.p, src.ix.sa, sa_bytes); memcpy(d_pac.p, src.ix.pac, pac_bytes);
#endif
    ix.bwt = d_bwt.p; ix.sa = d_sa.p; ix.pac = d_pac.p;
    return common_init(h, o, rg_id, c);
  }

  int init_device(const B2IndexHost& h, const void* bwt, const void* sa, const void* pac, const B2Opt& o,
                  const char* rg_id, const B2EngineConfig& c) {
    owns_index = false;
    bwt_bytes = (h.bwt_words ? h.bwt_words : h.bwt.size()) * 4; sa_bytes = (h.n_sa ? h.n_sa : h.sa.size()) * 8;
    pac_bytes = (size_t)(h.l_pac / 4 + 1);
    ix.bwt = (const uint32_t*)bwt; ix.sa = (const uint64_t*)sa; ix.pac = (const uint8_t*)pac;
    return common_init(h, o, rg_id, c);
  }

  ~B2EngineImpl() {
    for (auto& w : workers) {
      w.timer.destroy();
      for (DBuf<uint8_t>* b : {&w.seq, &w.has_qual, &w.scratch}) b->release();
      w.qual.release(); w.names.release(); w.comments.release(); w.seq_off.release(); w.name_off.release();
      w.com_off.release(); w.l_seq.release(); w.l_name.release(); w.l_com.release(); w.intv.release();
      w.seed_scratch.release(); w.seed_spill.release(); w.n_intv.release(); w.n_tasks.release(); w.seed_off.release(); w.raw.release();
      w.cseeds.release(); w.raw_rid.release(); w.chains.release(); w.n_chain.release(); w.reg_cap.release();
      w.reg_off.release(); w.regs.release(); w.n_regs.release(); w.pe_dir.release(); w.pe_is.release();
      w.pairtab.release(); w.slots.release(); w.sam_out.release(); w.sam_len.release(); w.sam_off.release();
      w.flags.release(); w.counters.release(); w.dbg.release(); w.fin_heavy.release(); w.fin_list.release(); w.sa_task_r.release(); w.sa_task_iv.release(); w.big.release(); w.dbg_seed.release();
      w.chain_off.release(); w.cand.release(); w.cand_ok.release();
      w.resc_cnt.release(); w.resc_off.release(); w.resc_tasks.release(); w.resc_res.release();
      w.xext_tasks.release(); w.xext_res.release(); w.xext_key.release(); w.xext_hist.release(); w.xext_perm.release(); w.xext_hoff.release();
      w.galn_cnt.release(); w.galn_table.release(); w.galn_off.release(); w.galn_tasks.release(); w.galn_res.release();
      for (void* q : w.pin_sams) if (q) b2_host_free(q);
      if (w.pin_in) b2_host_free(w.pin_in);
      w.raw_in.release(); w.com_cnt.release();
#if !defined(B2_HOSTSIM)
      if (w.stream) cudaStreamDestroy(w.stream);
#endif
    }
    if (owns_index) { d_bwt.release(); d_sa.release(); d_pac.release(); }
    d_anns.release(); d_names.release(); d_logtab.release(); d_rg.release();
  }

  // the CUDA device is a per-thread setting: every entry point that may run on a caller's thread selects it
  void bind_device() {
#if !defined(B2_HOSTSIM)
    cudaSetDevice(cfg.device);
#endif
  }

  int check_cuda(const char* where) {
#if !defined(B2_HOSTSIM)
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(where) + ": " + cudaGetErrorString(e));
#endif
    (void)where;
    return 0;
  }

  int upload(Worker& w, const B2HostBatch& b) {
    bind_device();
    const int n = b.n_reads;
    w.timer.reset();
    w.times = B2StageTimes();
    int t0 = w.timer.mark();
    if (w.seq.ensure(b.seq.size() + 8) || w.qual.ensure(b.seq.size() + 8) || w.has_qual.ensure(n + 1) ||
        w.names.ensure(b.names.size() + 8) || w.comments.ensure(b.comments.size() + 8) || w.seq_off.ensure(n + 1) ||
        w.name_off.ensure(n + 1) || w.com_off.ensure(n + 1) || w.l_seq.ensure(n + 1) || w.l_name.ensure(n + 1) ||
        w.l_com.ensure(n + 1))
      return fail("device allocation failed (read batch)");
    b2_h2d(w.seq.p, b.seq.data(), b.seq.size(), w.stream);
    if (!b.qual.empty()) b2_h2d(w.qual.p, b.qual.data(), b.qual.size(), w.stream);
    b2_h2d(w.has_qual.p, b.has_qual.data(), n, w.stream);
    b2_h2d(w.names.p, b.names.data(), b.names.size(), w.stream);
    if (!b.comments.empty()) b2_h2d(w.comments.p, b.comments.data(), b.comments.size(), w.stream);
    b2_h2d(w.seq_off.p, b.seq_off.data(), n * sizeof(int64_t), w.stream);
    b2_h2d(w.name_off.p, b.name_off.data(), n * sizeof(int64_t), w.stream);
    b2_h2d(w.com_off.p, b.com_off.data(), n * sizeof(int64_t), w.stream);
    b2_h2d(w.l_seq.p, b.l_seq.data(), n * sizeof(int32_t), w.stream);
    b2_h2d(w.l_name.p, b.l_name.data(), n * sizeof(int32_t), w.stream);
    b2_h2d(w.l_com.p, b.l_com.data(), n * sizeof(int32_t), w.stream);
    b2_g_h2d_bytes += b.seq.size() + b.qual.size() + n + b.names.size() + b.comments.size() + (size_t)n * (3 * 8 + 3 * 4);
    w.h2d_bytes += b.seq.size() + b.qual.size() + n + b.names.size() + b.comments.size() + (size_t)n * (3 * 8 + 3 * 4);
    int t1 = w.timer.mark();
    b2_sync(w.stream);
    w.times.h2d = w.timer.ms(t0, t1);
    w.n_reads = n; w.max_len = b.max_len; w.n_processed = b.n_processed;
    w.resident = true;
    return check_cuda("upload");
  }

  // One batch given as raw FASTQ bytes (B2RawInput): the byte range its records span in each mapped file goes to a
  // pinned staging buffer together with the record descriptors, one H2D copy, then the ingest kernels build the
  // packed batch on the device.  No synchronisation: everything is queued on the worker's stream ahead of run().
  int upload_raw(Worker& w, const B2RawInput& in) {
    bind_device();
    const int n = in.n_reads;
    w.timer.reset();
    w.times = B2StageTimes();
    w.ev_in0 = w.timer.mark();
    size_t lo[2] = {0, 0}, sz[2] = {0, 0}, tot_name = 0, tot_com = 0;
    for (int f = 0; f < in.n_files; ++f) {
      const B2Rec* r = in.recs[f];
      const size_t m = in.n_recs[f];
      if (m == 0) return fail("raw batch without records");
      lo[f] = r[0].off;
      sz[f] = (size_t)(r[m - 1].off + r[m - 1].qual_rel + r[m - 1].l_seq - lo[f]);
      for (size_t i = 0; i < m; ++i) { tot_name += r[i].l_name; tot_com += r[i].l_hdr - r[i].l_name; }
    }
    if ((size_t)n != in.n_recs[0] + in.n_recs[1]) return fail("raw batch: record count mismatch");
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o1 = up16(sz[0]), orec0 = up16(o1 + sz[1]), orec1 = orec0 + in.n_recs[0] * sizeof(B2Rec);
    const size_t total = orec1 + in.n_recs[1] * sizeof(B2Rec);
    if (total > w.pin_in_cap) {
      if (w.pin_in) b2_host_free(w.pin_in);
      w.pin_in = nullptr;
      w.pin_in_cap = total + total / 2 + ((size_t)1 << 20);
      if (b2_host_alloc(&w.pin_in, w.pin_in_cap)) { w.pin_in_cap = 0; return fail("pinned host allocation failed"); }
    }
    uint8_t* st = (uint8_t*)w.pin_in;
    const auto t_st0 = std::chrono::steady_clock::now();
    memcpy(st, in.base[0] + lo[0], sz[0]);
    if (in.n_files == 2) memcpy(st + o1, in.base[1] + lo[1], sz[1]);
    memcpy(st + orec0, in.recs[0], in.n_recs[0] * sizeof(B2Rec));
    if (in.n_files == 2) memcpy(st + orec1, in.recs[1], in.n_recs[1] * sizeof(B2Rec));
    w.stage_host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_st0).count();
    const size_t nb = (size_t)in.n_bases;
    if (w.raw_in.ensure(total + 16) || w.seq.ensure(nb + 8) || w.qual.ensure(nb + 8) || w.has_qual.ensure(n + 1) ||
        w.names.ensure(tot_name + 8) || w.comments.ensure((in.copy_comment ? tot_com : 0) + 8) || w.seq_off.ensure(n + 2) ||
        w.name_off.ensure(n + 2) || w.com_off.ensure(n + 2) || w.l_seq.ensure(n + 1) || w.l_name.ensure(n + 1) ||
        w.l_com.ensure(n + 1) || w.com_cnt.ensure(n + 1))
      return fail("device allocation failed (read batch)");
    b2_h2d(w.raw_in.p, st, total, w.stream);
    b2_g_h2d_bytes += total;
    w.h2d_bytes += total;
    B2Ingest g;
    memset(&g, 0, sizeof(g));
    g.raw = w.raw_in.p;
    g.recs[0] = (const B2Rec*)(w.raw_in.p + orec0); g.recs[1] = (const B2Rec*)(w.raw_in.p + orec1);
    g.adj[0] = -(int64_t)lo[0]; g.adj[1] = (int64_t)o1 - (int64_t)lo[1];
    g.n_files = in.n_files; g.n_reads = n; g.copy_comment = in.copy_comment;
    g.seq = w.seq.p; g.qual = w.qual.p; g.has_qual = w.has_qual.p; g.names = w.names.p; g.comments = w.comments.p;
    g.seq_off = w.seq_off.p; g.name_off = w.name_off.p; g.com_off = w.com_off.p;
    g.l_seq = w.l_seq.p; g.l_name = w.l_name.p; g.l_com = w.l_com.p; g.com_cnt = w.com_cnt.p;
    const int grid = std::min((n + 127) / 128, 148 * 8);
    B2_LAUNCH(k_ingest_len, grid, 128, w.stream, g);
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.l_seq.p, w.seq_off.p, n);
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.l_name.p, w.name_off.p, n);
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.com_cnt.p, w.com_off.p, n);
    B2_LAUNCH(k_ingest_copy, std::min((4 * n + 127) / 128, 148 * 16), 128, w.stream, g);
    w.ev_in1 = w.timer.mark();
    w.n_reads = n; w.max_len = in.max_len; w.n_processed = in.n_processed;
    w.resident = true;
    return check_cuda("upload_raw");
  }

  int read_flags(Worker& w, int* f) {
    b2_d2h(f, w.flags.p, sizeof(int), w.stream);
    if (b2_sync(w.stream)) {
#if !defined(B2_HOSTSIM)
      return fail(std::string("device synchronisation failed: ") + cudaGetErrorString(cudaGetLastError()));
#else
      return fail("device synchronisation failed");
#endif
    }
    return check_cuda("kernel");
  }

  // Runs the pipeline over reads [first, first+count) of the batch resident on `src` using the stream and
  // buffers of `w` (src == w for the ordinary one-batch call).
  // pe_mode: -1 = as the options say; 0 / 1 = this batch single-end / paired-end (smart pairing runs both kinds)
  int run(Worker& w, const B2Pes* pes0, std::string* sam, Worker* srcp = nullptr, int first = 0, int count = -1,
          bool view = false, int pe_mode = -1, std::vector<int32_t>* item_len = nullptr) {
    w.last_sam_bytes = 0;
    bind_device();
    Worker& src = srcp ? *srcp : w;
    if (!src.resident) return fail("no resident batch");
    if (count < 0) count = src.n_reads - first;
    if (first < 0 || first + count > src.n_reads) return fail("resident range out of bounds");
    const int n = count;
    const bool pe = pe_mode >= 0 ? pe_mode != 0 : (opt.flag & B2_F_PE) != 0;
    if (pe && (n & 1)) return fail("paired-end batch with an odd number of reads");
    const bool ingest = w.ev_in0 >= 0;   // upload_raw queued the ingest ahead of this call: its events are still pending
    if (!ingest) w.timer.reset();
    float t_h2d = w.times.h2d;
    w.times = B2StageTimes();
    w.times.h2d = t_h2d;
    if (ingest) { w.times.launches = 5; w.times.stage_host = w.stage_host_ms; }
    const int ev_in0 = w.ev_in0, ev_in1 = w.ev_in1;
    w.ev_in0 = w.ev_in1 = -1;
    if (n == 0) return 0;
    B2Ctx c;
    memset(&c, 0, sizeof(c));
    c.opt = opt; c.ix = ix; c.fmt = fmt;
    c.opt.flag = (c.opt.flag & ~(B2_F_PE | B2_F_SMARTPE)) | (pe ? B2_F_PE : 0);
    c.tb.logtab = d_logtab.p; c.tb.n_log = n_log;
    c.n_processed = src.n_processed + first; c.max_len = src.max_len;
    w.max_len = src.max_len;
    c.b.n_reads = n; c.b.seq = src.seq.p; c.b.qual = src.qual.p; c.b.has_qual = src.has_qual.p + first; c.b.names = src.names.p;
    c.b.comments = src.comments.p; c.b.seq_off = src.seq_off.p + first; c.b.l_seq = src.l_seq.p + first;
    c.b.name_off = src.name_off.p + first; c.b.l_name = src.l_name.p + first; c.b.com_off = src.com_off.p + first;
    c.b.l_com = src.l_com.p + first;
    w.n_reads_run = n;
    c.flags = w.flags.p;
    c.counters = w.counters.p;
    c.work = w.flags.p + 1;
    c.dbg = nullptr;
    // k_chain uses one lane per read, k_extend/k_final one B2_TASK_GROUP-lane group per read/pair; the scratch
    // region is indexed by lane resp. group, so size it for the larger count
    // the scratch area is [lanes of the launch] x [bytes per lane]: when one read needs far more than the common
    // case (long reads, wide bands: the traceback matrix of a global alignment is band x length bytes), the stage is
    // re-run with FEWER lanes rather than with that size for every lane -- the total stays within max_scratch_bytes
    int lane_grid = cfg.lane_grid;
    bool overflowed = false;
    // sizes another stream/buffer set of this engine had to grow to (an outlier read or pair) are taken over at the next
    // batch boundary: every set meets such input sooner or later, and a growth in mid-batch costs a re-run of the stage
    // plus a device-wide allocation
    w.scratch_per_lane = std::max(w.scratch_per_lane, grown_scratch.load());
    w.big_slot = std::max(w.big_slot, grown_big_slot.load());
    w.cap_intv = std::max(w.cap_intv, grown_cap_intv.load());
    if (w.slot_size) w.slot_size = std::max(w.slot_size, grown_slot.load());
    const size_t scratch_in = w.scratch_per_lane, big_slot_in = w.big_slot;
    const int cap_intv_in = w.cap_intv;
    if (w.scratch_per_lane > cfg.scratch_per_lane && w.clean_batches >= w.decay_wait &&
        cfg.max_scratch_bytes / w.scratch_per_lane / (size_t)cfg.lane_block < (size_t)cfg.lane_grid) {
      // an outlier batch grew the per-lane size so far that the stages now run with fewer lanes than configured: try
      // the smaller size again (the allocation itself is kept: freeing device memory stalls every stream)
      w.scratch_per_lane = std::max(cfg.scratch_per_lane, w.scratch_per_lane / 2);
      w.tried_decay = true;
      grown_scratch.store(w.scratch_per_lane);
    }
    int flags = 0;
    int ev_begin = w.timer.mark();

    // ---- seeding ----
    // one read per lane; the lane count is what fits on the chip at once (shared-memory interval stacks), the
    // lanes pull reads from a cursor
    if (w.max_len >= (1 << 24)) return fail("read of 16 777 216 bases or more (the seeding kernel keeps query positions in 24 bits)");
    const int seed_block = cfg.seed_block;
    int seed_QW = std::min((w.max_len + 7) / 8, 32), seed_RS = 4, seed_E;
    {  // split the per-lane share of the SM's shared memory: bases, candidates, the rest for the interval stack
      const int per_lane_words = (int)((227 * 1024 / cfg.seed_blocks_per_sm - 1024) / seed_block / 4);
      seed_E = (per_lane_words - seed_QW - 2 * seed_RS) / 3;
      if (cfg.seed_smem_entries > 0) seed_E = cfg.seed_smem_entries;
      if (seed_E < 4) seed_E = 4;
    }
    const size_t seed_smem = (size_t)seed_block * (3 * seed_E + seed_QW + 2 * seed_RS) * sizeof(uint32_t);
    const size_t seed3_smem = (size_t)seed_block * seed_QW * sizeof(uint32_t);
    int seed_grid = std::min({(n + seed_block - 1) / seed_block, 148 * cfg.seed_blocks_per_sm, cfg.seed_grid});
    int seed3_grid = std::min({(n + seed_block - 1) / seed_block, 148 * cfg.seed3_blocks_per_sm, cfg.seed_grid});
    {  // bound the spill area for very long reads
      const size_t per_lane = 5 * sizeof(uint32_t) * (size_t)(w.max_len + 1);
      const size_t max_lanes = std::max<size_t>(((size_t)2 << 30) / per_lane, (size_t)seed_block);
      seed_grid = (int)std::min<size_t>((size_t)seed_grid, max_lanes / seed_block);
      seed3_grid = (int)std::min<size_t>((size_t)seed3_grid, max_lanes / seed_block);
      if (seed_grid < 1) seed_grid = 1;
      if (seed3_grid < 1) seed3_grid = 1;
    }
    const size_t seed_lanes = (size_t)std::max(seed_grid, seed3_grid) * seed_block;
#if defined(B2_HOSTSIM)
    if (w.seed_scratch.ensure(seed_lanes * (size_t)(w.max_len + 1))) return fail("device allocation failed (seeding)");
    (void)seed_smem; (void)seed3_smem;
#else
    if (w.seed_spill.ensure(seed_lanes * 5 * (size_t)(w.max_len + 1) + 8)) return fail("device allocation failed (seeding)");
    if (seed_smem > (size_t)227 * 1024) return fail("seeding shared-memory configuration exceeds 227 KB per block");
#endif
    if (w.n_intv.ensure(n + 1) || w.n_tasks.ensure(n + 1) || w.seed_off.ensure(n + 2)) return fail("device allocation failed (seeding)");
    for (;;) {
      if (w.intv.ensure((size_t)n * w.cap_intv)) return fail("device allocation failed (intervals)");
      c.intv = w.intv.p; c.cap_intv = w.cap_intv; c.n_intv = w.n_intv.p; c.n_tasks = w.n_tasks.p;
      c.seed_scratch = w.seed_scratch.p; c.seed_spill = w.seed_spill.p;
      c.seed_E = seed_E; c.seed_QW = seed_QW; c.seed_RS = seed_RS;
      b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);
      c.dbg_seed = nullptr;
      if (getenv("B200BWA_DEBUG_TIMES")) {
        if (w.dbg_seed.ensure(n + 1)) return fail("device allocation failed");
        b2_dev_memset(w.dbg_seed.p, 0, (size_t)n * sizeof(int32_t), w.stream);
        c.dbg_seed = w.dbg_seed.p;
      }
      B2_LAUNCH_SMEM(k_seed, seed_grid, seed_block, seed_smem, w.stream, c);
      B2_LAUNCH_SMEM(k_seed3, seed3_grid, seed_block, seed3_smem, w.stream, c);
      B2_LAUNCH(k_seed_fin, (n + 127) / 128, 128, w.stream, c);
      w.times.launches += 3;
      if (read_flags(w, &flags)) return 1;
      if (!(flags & B2_FLAG_OVF_INTV)) break;
      // a long read seeded with a short -k has more intervals than bases (re-seeding, bwamem.c:131-144): past 8 192
      // entries per read the table keeps doubling for as long as the whole batch's stays within 4 GB
      if (w.cap_intv >= 8192 && (size_t)n * (size_t)w.cap_intv * 2 * sizeof(B2Intv) > ((size_t)4 << 30))
        return fail("seed interval capacity exceeded");
      w.cap_intv *= 2;
    }
    if (c.dbg_seed) {   // extension steps per read: how uneven the lanes' work is (one read per lane)
      std::vector<int32_t> h(n);
      b2_d2h(h.data(), w.dbg_seed.p, (size_t)n * sizeof(int32_t), w.stream);
      b2_sync(w.stream);
      double sum = 0, wmax = 0, bmax = 0;
      for (int i = 0; i < n; ++i) sum += h[i];
      for (int i = 0; i + 32 <= n; i += 32) wmax += *std::max_element(h.begin() + i, h.begin() + i + 32);
      for (int i = 0; i + 128 <= n; i += 128) bmax += *std::max_element(h.begin() + i, h.begin() + i + 128);
      std::vector<int32_t> srt(h);
      std::sort(srt.begin(), srt.end());
      auto pct = [&](double f) { return srt[std::min<size_t>((size_t)n - 1, (size_t)(f * n))]; };
      fprintf(stderr, "[D::seed] steps per read (passes 1+2): mean %.1f p10 %d p50 %d p90 %d p99 %d p99.9 %d max %d; mean / (mean of the maxima of 32 consecutive reads) %.3f, of 128 %.3f\n",
              sum / n, pct(.1), pct(.5), pct(.9), pct(.99), pct(.999), srt[n - 1], sum / n / (wmax / (n / 32)), sum / n / (bmax / (n / 128)));
    }
    int ev_seed = w.timer.mark();
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.n_tasks.p, w.seed_off.p, n);
    ++w.times.launches;
    int64_t S = 0;
    b2_d2h(&S, w.seed_off.p + n, sizeof(int64_t), w.stream);
    if (b2_sync(w.stream)) return fail("device synchronisation failed");
    c.seed_off = w.seed_off.p; c.n_seed_tasks = S;
    w.times.n_seed_tasks = S;

    // ---- SA lookup ----
    if (w.raw.ensure(S + 1) || w.raw_rid.ensure(S + 1) || w.chains.ensure(S + 1) || w.cseeds.ensure(S + 1) ||
        w.n_chain.ensure(n + 1) || w.reg_cap.ensure(n + 1) || w.reg_off.ensure(n + 2) || w.n_regs.ensure(n + 1))
      return fail("device allocation failed (seeds)");
    c.raw = w.raw.p; c.raw_rid = w.raw_rid.p; c.chains = w.chains.p; c.cseeds = w.cseeds.p;
    c.n_chain = w.n_chain.p; c.reg_cap = w.reg_cap.p; c.n_regs = w.n_regs.p;
    c.sa_task_r = nullptr; c.sa_task_iv = nullptr;
    if (S > 0 && !cfg.no_sa_plan && opt.max_occ <= 0xffff && w.cap_intv <= 0xffff) {
      if (w.sa_task_r.ensure(S + 1) || w.sa_task_iv.ensure(S + 1)) return fail("device allocation failed (seeds)");
      c.sa_task_r = w.sa_task_r.p; c.sa_task_iv = w.sa_task_iv.p;
      B2_LAUNCH(k_sa_plan, (n + 127) / 128, 128, w.stream, c);
      ++w.times.launches;
    }
    if (S > 0) {
      int grid = (int)std::min<int64_t>((S + 127) / 128, 148 * 16);   // resident lanes only (all 64 warps of an SM: the walk is a chain of dependent fetches); each takes lookup after lookup
      b2_dev_memset(w.flags.p + 3, 0, sizeof(int), w.stream);   // the lookups are dealt from this cursor (c.work[2])
      if (c.sa_task_r) B2_LAUNCH(k_sa, grid, 128, w.stream, c);
      else B2_LAUNCH(k_sa_search, grid, 128, w.stream, c);
      B2_LAUNCH(k_sa_rid, (int)std::min<int64_t>((S + 255) / 256, 148 * 16), 256, w.stream, c);
      w.times.launches += 2;
    }
    int ev_sa = w.timer.mark();

    // ---- lane-per-read stages with retry on scratch overflow ----
    auto ensure_scratch = [&]() -> int {
      const size_t fit = cfg.max_scratch_bytes / w.scratch_per_lane / (size_t)cfg.lane_block;
      if (fit < 1) return fail("per-lane scratch limit exceeded (a single read needs more than B200BWA_MAX_SCRATCH)");
      lane_grid = (int)std::min<size_t>((size_t)cfg.lane_grid, fit);
      if (w.scratch.ensure((size_t)lane_grid * cfg.lane_block * w.scratch_per_lane, true)) return fail("device allocation failed (scratch)");
      c.scratch = w.scratch.p; c.scratch_per_lane = w.scratch_per_lane;
      c.big_pool = nullptr; c.big_n = 0; c.big_slot = 0; c.big_count = w.flags.p + 2;
      if (cfg.big_pool_bytes > 0) {   // outlier pool (see B2Ctx::big_pool)
        if (w.big_slot > cfg.max_scratch_bytes) return fail("per-lane scratch limit exceeded (a single read needs more than B200BWA_MAX_SCRATCH)");
        const size_t bytes = std::max(std::min(cfg.big_pool_bytes, cfg.max_scratch_bytes), w.big_slot);
        if (w.big.ensure(bytes, true)) return fail("device allocation failed (scratch pool)");
        c.big_pool = w.big.p; c.big_slot = w.big_slot; c.big_n = (int)std::min<size_t>(w.big.cap / w.big_slot, (size_t)1 << 30);
      }
      return 0;
    };
    // a stage flagged an overflow: inside a pool slot -> larger slots; more outliers than slots (or no pool) -> larger
    // regions for every lane (then fewer lanes, within max_scratch_bytes)
    auto grow_scratch = [&](int fl) {
      overflowed = true;
      if (fl & B2_FLAG_OVF_BIGN) w.scratch_per_lane *= 2;
      if (fl & B2_FLAG_OVF_SCRATCH) { if (c.big_pool) w.big_slot *= 2; else w.scratch_per_lane *= 2; }
    };
    const int OVF_ANY = B2_FLAG_OVF_SCRATCH | B2_FLAG_OVF_BIGN;
    for (;;) {
      if (ensure_scratch()) return 1;
      b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
      B2_LAUNCH(k_chain, lane_grid, cfg.lane_block, w.stream, c);
      ++w.times.launches;
      if (read_flags(w, &flags)) return 1;
      if (!(flags & OVF_ANY)) break;
      grow_scratch(flags);
    }
    {  // mem_flt_chained_seeds acts only on long reads (5.5 ln l <= 0.05 l: l >= 725 with the default options)
      const double min_l = opt.min_chain_weight ? (double)(1.1f * (float)opt.min_chain_weight) : 5.5 * log((double)w.max_len);
      if (!(min_l > (double)(0.05f * (float)w.max_len)) && !getenv("B200BWA_DEBUG_NO_SEED_FILTER")) {  // (the switch exists to show the filter matters)
        for (;;) {
          if (ensure_scratch()) return 1;
          b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
          B2_LAUNCH(k_flt_seeds, lane_grid, cfg.lane_block, w.stream, c);
          ++w.times.launches;
          if (read_flags(w, &flags)) return 1;
          if (flags & B2_FLAG_ERR_TABLE) return fail("log table index out of range (read too long)");
          if (!(flags & OVF_ANY)) break;
          grow_scratch(flags);
        }
      }
    }
    int ev_chain = w.timer.mark();
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.reg_cap.p, w.reg_off.p, n);
    ++w.times.launches;
    int64_t R = 0;
    b2_d2h(&R, w.reg_off.p + n, sizeof(int64_t), w.stream);
    if (b2_sync(w.stream)) return fail("device synchronisation failed");
    if (w.regs.ensure(R + 1)) return fail("device allocation failed (regions)");
    c.reg_off = w.reg_off.p; c.regs = w.regs.p;
    w.times.n_regs = R;
    // speculative per-chain extension tasks, then the serial per-read pass
    if (w.chain_off.ensure(n + 2) || w.cand.ensure(S + 1) || w.cand_ok.ensure(S + 1)) return fail("device allocation failed (extension)");
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.n_chain.p, w.chain_off.p, n);
    ++w.times.launches;
    int64_t TC = 0;
    b2_d2h(&TC, w.chain_off.p + n, sizeof(int64_t), w.stream);
    if (b2_sync(w.stream)) return fail("device synchronisation failed");
    c.chain_off = w.chain_off.p; c.n_chain_tasks = TC; c.cand = w.cand.p; c.cand_ok = w.cand_ok.p;
    // inter-task extension of every chain's best seed, left then right, ahead of the group-cooperative kernel
    int ev_x[4] = {-1, -1, -1, -1}, ev_r[2] = {-1, -1}, ev_g[2] = {-1, -1};   // CUDA events around the three DP kernels proper
    c.xext_res = nullptr; c.xext_tasks = nullptr;
    if (TC > 0 && !cfg.no_xext && b2_mat_simple(opt.mat) && S < (1 << 29)) {
      c.xext_qcap = w.max_len; c.xext_tcap = w.max_len + 2 * opt.w + 64;
      // {h, e} cells of 8 + 8 bits when no score of the batch can reach 256 (an extension's scores are bounded by
      // read length x match score), 16 + 16 bits otherwise
      c.xext_cell8 = (opt.a * w.max_len <= 250 && !getenv("B200BWA_XEXT_CELL32")) ? 1 : 0;
      const size_t ehw = (((size_t)c.xext_qcap + 2) * (c.xext_cell8 ? 2 : 4) + 3) / 4;
      const size_t words = (size_t)32 * (ehw + c.xext_qcap / 8 + 1 + c.xext_tcap / 8 + 1);
      const size_t smem = words * sizeof(uint32_t);
      if (smem <= (size_t)227 * 1024) {
        const int bins = c.xext_qcap + 1;
        if (w.xext_tasks.ensure(TC + 1) || w.xext_res.ensure(2 * (size_t)S + 2) || w.xext_key.ensure(TC + 1) || w.xext_perm.ensure(TC + 1) ||
            w.xext_hist.ensure(2 * (size_t)bins + 2) || w.xext_hoff.ensure((size_t)bins + 2))
          return fail("device allocation failed (extension tasks)");
        c.xext_tasks = w.xext_tasks.p; c.xext_res = w.xext_res.p; c.xext_key = w.xext_key.p; c.xext_perm = w.xext_perm.p;
        c.xext_hist = w.xext_hist.p; c.xext_cur = w.xext_hist.p + bins + 1; c.xext_hoff = w.xext_hoff.p;
        b2_dev_memset(w.xext_res.p, 0xff, (2 * (size_t)S + 2) * sizeof(B2XextRes), w.stream);  // w = -1: no result
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(24, ((size_t)227 * 1024) / smem));
        const int xgrid = (int)std::min<int64_t>((TC + 31) / 32, (int64_t)148 * per_sm);
        const int pgrid = (int)std::min<int64_t>((TC + 127) / 128, 148 * 8);
        for (int side = 0; side < 2; ++side) {
          b2_dev_memset(w.xext_hist.p, 0, (2 * (size_t)bins + 2) * sizeof(int32_t), w.stream);
          B2_LAUNCH(k_xext_plan, pgrid, 128, w.stream, c, side);
          B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.xext_hist.p, w.xext_hoff.p, bins);
          B2_LAUNCH(k_xext_scatter, pgrid, 128, w.stream, c);
          ev_x[2 * side] = w.timer.mark();
          B2_LAUNCH_SMEM(k_xext_run, xgrid, 32, smem, w.stream, c);
          ev_x[2 * side + 1] = w.timer.mark();
          w.times.launches += 4;
        }
      }
    }
    for (;;) {
      if (ensure_scratch()) return 1;
      if (TC > 0 && cfg.no_ext_chain) b2_dev_memset(w.cand_ok.p, 0, (size_t)S + 1, w.stream);
      if (TC > 0 && !cfg.no_ext_chain) {
        b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
        const int64_t blocks_needed = (TC * B2_G_EXT + cfg.lane_block - 1) / cfg.lane_block;
        B2_LAUNCH(k_ext_chain, (int)std::min<int64_t>(blocks_needed, lane_grid), cfg.lane_block, w.stream, c);
        ++w.times.launches;
        if (read_flags(w, &flags)) return 1;
        if (flags & OVF_ANY) { grow_scratch(flags); continue; }
      }
      b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
      B2_LAUNCH(k_extend, lane_grid, cfg.lane_block, w.stream, c);
      ++w.times.launches;
      if (read_flags(w, &flags)) return 1;
      if (!(flags & OVF_ANY)) break;
      grow_scratch(flags);
    }
    int ev_ext = w.timer.mark();

    // ---- insert-size statistics (batch-wide barrier) ----
    const int n_items = pe ? n >> 1 : n;
    int bad_pair = 0x7f7f7f7f;
    if (pe) {
      b2_dev_memset(w.flags.p + 6, 0x7f, sizeof(int), w.stream);
      B2_LAUNCH(k_pair_names, (n_items + 255) / 256, 256, w.stream, c, w.flags.p + 6);
      ++w.times.launches;
      b2_d2h(&bad_pair, w.flags.p + 6, sizeof(int), w.stream);   // (read after the stage's closing synchronisation below)
      if (pes0) memcpy(c.pes, pes0, sizeof(B2Pes) * 4);
      else {
        if (w.pe_dir.ensure(n_items + 1) || w.pe_is.ensure(n_items + 1)) return fail("device allocation failed");
        c.pe_dir = w.pe_dir.p; c.pe_is = w.pe_is.p;
        B2_LAUNCH(k_pestat, (n_items + 127) / 128, 128, w.stream, c);
        ++w.times.launches;
        w.h_dir.resize(n_items); w.h_is.resize(n_items);
        b2_d2h(w.h_dir.data(), w.pe_dir.p, n_items, w.stream);
        b2_d2h(w.h_is.data(), w.pe_is.p, n_items * sizeof(int64_t), w.stream);
        if (b2_sync(w.stream)) return fail("device synchronisation failed");
        b2_pestat_host(opt, n_items, w.h_dir.data(), w.h_is.data(), c.pes, cfg.verbose);
      }
      // pairing table: .721*log(2*erfc(|ns|/sqrt2))*a for every admissible insert size
      std::vector<double> tab;
      size_t offs[4] = {0, 0, 0, 0};
      for (int d = 0; d < 4; ++d) {
        c.tb.pair_n[d] = 0;
        if (c.pes[d].failed || c.pes[d].high < c.pes[d].low) continue;
        offs[d] = tab.size();
        long cnt = (long)c.pes[d].high - c.pes[d].low + 1;
        if (cnt > (1 << 24)) return fail("insert-size range too wide for the pairing table");
        c.tb.pair_n[d] = (int)cnt;
        for (long k = 0; k < cnt; ++k) {
          int64_t dist = (int64_t)c.pes[d].low + k;
          double ns = (dist - c.pes[d].avg) / c.pes[d].std;
          tab.push_back(.721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt.a);
        }
      }
      if (w.pairtab.ensure(tab.size() + 1)) return fail("device allocation failed");
      if (!tab.empty()) b2_h2d(w.pairtab.p, tab.data(), tab.size() * sizeof(double), w.stream);
      for (int d = 0; d < 4; ++d) c.tb.pairtab[d] = w.pairtab.p + offs[d];
      b2_sync(w.stream);
      if (bad_pair != 0x7f7f7f7f) {   // as the reference: the run ends here (it exits; this call returns 1)
        std::string nm[2];
        for (int k = 0; k < 2; ++k) {
          int64_t off = 0; int32_t len = 0;
          b2_d2h(&off, c.b.name_off + 2 * bad_pair + k, sizeof(off), w.stream);
          b2_d2h(&len, c.b.l_name + 2 * bad_pair + k, sizeof(len), w.stream);
          b2_sync(w.stream);
          nm[k].resize((size_t)(len > 0 ? len : 0));
          if (len > 0) { b2_d2h(&nm[k][0], c.b.names + off, (size_t)len, w.stream); b2_sync(w.stream); }
        }
        return fail("[mem_sam_pe] paired reads have different names: \"" + nm[0] + "\", \"" + nm[1] + "\"");
      }
    }
    int ev_pes0 = w.timer.mark();
    // ---- mate-rescue attempts as independent SW tasks ------------------------------------------------
    c.resc_off = nullptr; c.n_resc = 0;
    if (pe && !(opt.flag & B2_F_NO_RESCUE) && !cfg.no_resc_plan) {
      if (w.resc_cnt.ensure(n_items + 1) || w.resc_off.ensure(n_items + 2)) return fail("device allocation failed");
      c.resc_cnt = w.resc_cnt.p;
      B2_LAUNCH(k_resc_plan, (n_items + 127) / 128, 128, w.stream, c, 0);
      B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.resc_cnt.p, w.resc_off.p, n_items);
      w.times.launches += 2;
      int64_t T = 0;
      b2_d2h(&T, w.resc_off.p + n_items, sizeof(int64_t), w.stream);
      if (b2_sync(w.stream)) return fail("device synchronisation failed");
      if (w.resc_tasks.ensure(T + 1) || w.resc_res.ensure(T + 1)) return fail("device allocation failed (rescue tasks)");
      c.resc_off = w.resc_off.p; c.resc_tasks = w.resc_tasks.p; c.resc_res = w.resc_res.p; c.n_resc = T;
      w.times.n_resc = T;
      if (T > 0) {
        B2_LAUNCH(k_resc_plan, (n_items + 127) / 128, 128, w.stream, c, 1);
        ++w.times.launches;
        for (;;) {
          if (ensure_scratch()) return 1;
          b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
          const int64_t groups_needed = (T * B2_G_RESC + cfg.lane_block - 1) / cfg.lane_block;
          const int grid = (int)std::min<int64_t>(groups_needed, lane_grid);
          ev_r[0] = w.timer.mark();
          B2_LAUNCH(k_resc_sw, grid, cfg.lane_block, w.stream, c);
          ev_r[1] = w.timer.mark();
          ++w.times.launches;
          if (read_flags(w, &flags)) return 1;
          if (!(flags & OVF_ANY)) break;
          grow_scratch(flags);
        }
      }
    }
    int ev_pes = w.timer.mark();

    // ---- end-point-constrained global alignments, one per lane, ahead of the finalisation (b2_galn.cuh) ----
    memset(&c.galn, 0, sizeof(c.galn));
    {
      const size_t half = ((size_t)w.max_len + 64 + 15) & ~(size_t)15;
      const size_t pool_need = 32 * 2 * half + ((size_t)w.max_len + 64) * B2_GALN_MAX_ZWORDS * 32 * 4;
      if (!cfg.no_galn && b2_mat_simple(opt.mat) && !ensure_scratch() && pool_need <= 32 * w.scratch_per_lane) {
        if (w.galn_cnt.ensure((size_t)B2_GALN_CLASSES * (n + 1)) || w.galn_off.ensure((size_t)B2_GALN_CLASSES * (n + 2)))
          return fail("device allocation failed (alignment plan)");
        c.galn_cnt = w.galn_cnt.p; c.galn_off = w.galn_off.p;
        B2_LAUNCH(k_galn_plan, (n + 127) / 128, 128, w.stream, c, 0);
        for (int k = 0; k < B2_GALN_CLASSES; ++k)
          B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)(w.galn_cnt.p + (size_t)k * (n + 1)), w.galn_off.p + (size_t)k * (n + 2), n);
        w.times.launches += 1 + B2_GALN_CLASSES;
        int64_t tot[B2_GALN_CLASSES], T = 0;
        for (int k = 0; k < B2_GALN_CLASSES; ++k) b2_d2h(&tot[k], w.galn_off.p + (size_t)k * (n + 2) + n, sizeof(int64_t), w.stream);
        if (b2_sync(w.stream)) return fail("device synchronisation failed");
        for (int k = 0; k < B2_GALN_CLASSES; ++k) { c.galn_base[k] = T; c.galn_n[k] = tot[k]; T += tot[k]; }
        if (getenv("B200BWA_DEBUG_GALN")) fprintf(stderr, "[D::galn] reads %d tasks %ld (%ld, %ld, %ld, %ld)\n", n, (long)T, (long)tot[0], (long)tot[1], (long)tot[2], (long)tot[3]);
        if (T > 0 && T < (1 << 29)) {
          size_t tsz = 2;
          while (tsz < (size_t)2 * T) tsz <<= 1;
          if (w.galn_tasks.ensure(T + 1) || w.galn_res.ensure(T + 1) || w.galn_table.ensure(tsz))
            return fail("device allocation failed (alignment tasks)");
          b2_dev_memset(w.galn_table.p, 0xff, tsz * sizeof(int32_t), w.stream);
          c.galn_tasks = w.galn_tasks.p; c.galn_res = w.galn_res.p; c.galn_table = w.galn_table.p;
          c.galn.tasks = w.galn_tasks.p; c.galn.res = w.galn_res.p; c.galn.table = w.galn_table.p;
          c.galn.mask = (uint32_t)(tsz - 1); c.galn.seq_base = c.b.seq;
          B2_LAUNCH(k_galn_plan, (n + 127) / 128, 128, w.stream, c, 1);
          ++w.times.launches;
          // one-warp blocks: a batch has only a few hundred warps of these tasks, spread them over all SMs
          const int blk = std::min(cfg.lane_block, 32);
          auto grid_for = [&](int64_t cnt) { return (int)std::min<int64_t>((cnt + blk - 1) / blk, (int64_t)lane_grid * cfg.lane_block / blk); };
          ev_g[0] = w.timer.mark();
          {  // one launch for the four classes: each gets blocks in proportion to its tasks, within the scratch lanes
            const int64_t budget = (int64_t)lane_grid * cfg.lane_block / blk;
            int64_t want[B2_GALN_CLASSES], sum = 0;
            int n_nonempty = 0;
            for (int k = 0; k < B2_GALN_CLASSES; ++k) { want[k] = (tot[k] + blk - 1) / blk; sum += want[k]; n_nonempty += want[k] > 0; }
            (void)grid_for;
            if (budget >= n_nonempty) {
              const int64_t spare = budget - n_nonempty, extra = sum - n_nonempty;
              int nb = 0;
              for (int k = 0; k < B2_GALN_CLASSES; ++k) {
                c.galn_blk[k] = nb;
                if (want[k] > 0) nb += (int)(1 + (extra <= spare ? want[k] - 1 : (want[k] - 1) * spare / extra));
              }
              c.galn_blk[B2_GALN_CLASSES] = nb;
              B2_LAUNCH(k_galn, nb, blk, w.stream, c);
              ++w.times.launches;
            } else {  // fewer scratch lanes than classes (the host checker build's single lane): one class after the other
              for (int k = 0; k < B2_GALN_CLASSES; ++k) {
                if (want[k] == 0) continue;
                const int g1 = (int)std::max<int64_t>(1, std::min<int64_t>(want[k], budget));
                for (int j = 0; j <= B2_GALN_CLASSES; ++j) c.galn_blk[j] = j <= k ? 0 : g1;
                B2_LAUNCH(k_galn, g1, blk, w.stream, c);
                ++w.times.launches;
              }
            }
          }
          ev_g[1] = w.timer.mark();
        }
      }
    }

    // ---- finalisation + SAM text ----
    if (w.sam_len.ensure(n_items + 1) || w.sam_off.ensure(n_items + 2)) return fail("device allocation failed");
    c.sam_len = w.sam_len.p;
    if (w.slot_size == 0) {
      int per_read = cfg.slot_per_read ? cfg.slot_per_read : 3 * w.max_len + 384;
      w.slot_size = std::max(pe ? 2 * per_read : per_read, grown_slot.load());
    }
    const int slot_in = w.slot_size;
    // pairs with many regions are listed once and finalised by their own launch, a warp each (see k_final_plan)
    const bool split = pe && cfg.final_heavy_regs > 0;
    c.fin_mode = 0;
    if (split) {
      if (w.fin_heavy.ensure(n_items + 1) || w.fin_list.ensure(n_items + 1)) return fail("device allocation failed");
      c.fin_heavy = w.fin_heavy.p; c.fin_list = w.fin_list.p; c.fin_heavy_min = cfg.final_heavy_regs;
      b2_dev_memset(w.flags.p + 4, 0, sizeof(int), w.stream);   // number of listed pairs (c.work[3])
      B2_LAUNCH(k_final_plan, (n_items + 127) / 128, 128, w.stream, c);
      ++w.times.launches;
    }
    for (;;) {
      if (ensure_scratch()) return 1;
      if (w.slots.ensure((size_t)n_items * w.slot_size)) return fail("device allocation failed (SAM slots)");
      c.slots = w.slots.p; c.slot_size = w.slot_size;
      if (getenv("B200BWA_DEBUG_TIMES")) {
        if (w.dbg.ensure((size_t)n_items * 4 + 16)) return fail("device allocation failed");
        b2_dev_memset(w.dbg.p, 0, ((size_t)n_items * 4 + 16) * sizeof(long long), w.stream);
        c.dbg = w.dbg.p;
      }
      b2_dev_memset(w.flags.p, 0, 3 * sizeof(int), w.stream);   // flags, task cursor, pool slots handed out
      if (split) {
        c.fin_mode = 2;
        c.fin_fast = 2048;
        B2_LAUNCH_SMEM(k_final_pe_heavy, std::min(lane_grid, 148 * B2_LB_FINAL), cfg.lane_block, (size_t)(cfg.lane_block / B2_G_HEAVY) * c.fin_fast, w.stream, c);
        b2_dev_memset(w.flags.p + 1, 0, sizeof(int), w.stream);   // the cursor; the flags word keeps what the first launch raised
        c.fin_mode = 1;
        ++w.times.launches;
      }
      c.fin_fast = cfg.final_fast_bytes;
      if (pe) B2_LAUNCH_SMEM(k_final_pe, lane_grid, cfg.lane_block, (size_t)(cfg.lane_block / B2_G_FINAL) * c.fin_fast, w.stream, c);
      else B2_LAUNCH(k_final_se, lane_grid, cfg.lane_block, w.stream, c);
      ++w.times.launches;
      if (read_flags(w, &flags)) return 1;
      if (flags & B2_FLAG_ERR_TABLE) return fail("log/pairing table index out of range");
      if (!(flags & (OVF_ANY | B2_FLAG_OVF_SLOT))) break;
      if (flags & OVF_ANY) grow_scratch(flags);
      if (flags & B2_FLAG_OVF_SLOT) {
        // (a long read printed with -a -Y repeats its full sequence in every record: past 16 MB per read the slots keep
        // doubling while the batch's stay within 8 GB)
        if (w.slot_size > (1 << 24) && (w.slot_size >= (1 << 29) || (size_t)n_items * (size_t)w.slot_size * 2 > ((size_t)8 << 30)))
          return fail("SAM slot limit exceeded");
        w.slot_size *= 2;
      }
    }
    int ev_fin = w.timer.mark();
    if (c.dbg) {
      std::vector<long long> h((size_t)n_items * 4 + 16);
      b2_d2h(h.data(), w.dbg.p, h.size() * sizeof(long long), w.stream);
      b2_sync(w.stream);
      {
        const long long* ph = h.data() + (size_t)n_items * 4;
        fprintf(stderr, "[D::final] phase cycles: load %lld rescue %lld mark_primary %lld pair+mapq %lld xa %lld reg2aln %lld aln2sam %lld reg2sam(unpaired) %lld\n",
                ph[7], ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6]);
        std::vector<long long> cyc(n_items);
        for (int i = 0; i < n_items; ++i) cyc[i] = h[4 * (size_t)i];
        std::sort(cyc.begin(), cyc.end());
        long long all = 0, top1 = 0;
        for (int i = 0; i < n_items; ++i) { all += cyc[i]; if (i >= n_items - n_items / 100) top1 += cyc[i]; }
        auto pct = [&](double f) { return cyc[std::min<size_t>((size_t)n_items - 1, (size_t)(f * n_items))]; };
        if (n_items > 0)
          fprintf(stderr, "[D::final] cycles per item: p10 %lld p50 %lld p90 %lld p99 %lld p99.9 %lld max %lld; slowest 1%% of the items hold %.1f%% of the cycles\n",
                  pct(.10), pct(.50), pct(.90), pct(.99), pct(.999), cyc[n_items - 1], all ? 100.0 * top1 / all : 0.);
      }
      std::vector<int> ord(n_items);
      for (int i = 0; i < n_items; ++i) ord[i] = i;
      std::sort(ord.begin(), ord.end(), [&](int a, int b) { return h[4 * a] > h[4 * b]; });
      long long tot = 0, totsw = 0, totgl = 0;
      for (int i = 0; i < n_items; ++i) { tot += h[4 * i]; totsw += h[4 * i + 1]; totgl += h[4 * i + 2]; }
      fprintf(stderr, "[D::final] items %d total cycles %lld sw_calls %lld glob_calls %lld; median cycles %lld\n", n_items, tot, totsw, totgl, h[4 * ord[n_items / 2]]);
      for (int k = 0; k < 12 && k < n_items; ++k) {
        int i = ord[k];
        fprintf(stderr, "[D::final]  item %d cycles %lld sw %lld glob %lld regs %lld/%lld\n", i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3] >> 32, h[4 * i + 3] & 0xffffffff);
      }
      // histogram by sw calls
      long long cyc_by[4] = {0, 0, 0, 0}, cnt_by[4] = {0, 0, 0, 0};
      for (int i = 0; i < n_items; ++i) { int b = h[4 * i + 1] == 0 ? 0 : h[4 * i + 1] <= 2 ? 1 : h[4 * i + 1] <= 8 ? 2 : 3; cyc_by[b] += h[4 * i]; ++cnt_by[b]; }
      for (int b = 0; b < 4; ++b) fprintf(stderr, "[D::final]  sw-bin %d: items %lld cycles %lld (avg %lld)\n", b, cnt_by[b], cyc_by[b], cnt_by[b] ? cyc_by[b] / cnt_by[b] : 0);
    }
    B2_LAUNCH(k_scan, 1, 1024, w.stream, (const int32_t*)w.sam_len.p, w.sam_off.p, n_items);
    ++w.times.launches;
    int64_t total = 0;
    b2_d2h(&total, w.sam_off.p + n_items, sizeof(int64_t), w.stream);
    if (b2_sync(w.stream)) return fail("device synchronisation failed");
    if (w.sam_out.ensure(total + 16)) return fail("device allocation failed (SAM out)");
    c.sam_off = w.sam_off.p; c.sam_out = w.sam_out.p;
    {
      int64_t work = (int64_t)n_items * ((w.slot_size + 15) / 16);
      int grid = (int)std::min<int64_t>((work + 255) / 256, 148 * 32);
      B2_LAUNCH(k_gather, grid, 256, w.stream, c, n_items);
      ++w.times.launches;
    }
    int ev_gat = w.timer.mark();
    w.times.sam_bytes = total;
    if (item_len) {
      item_len->resize(n_items);
      b2_d2h(item_len->data(), w.sam_len.p, (size_t)n_items * sizeof(int32_t), w.stream);
    }
    int ev_d2h = ev_gat;
    if (sam || view) {
      const int fl = w.sam_flip;
      w.sam_flip ^= 1;
      if ((size_t)total > w.pin_sam_caps[fl]) {
        if (w.pin_sams[fl]) b2_host_free(w.pin_sams[fl]);
        w.pin_sams[fl] = nullptr;
        w.pin_sam_caps[fl] = (size_t)total + (size_t)total / 2 + ((size_t)1 << 20);  // (cudaFreeHost/cudaMallocHost stall the device too)
        if (b2_host_alloc(&w.pin_sams[fl], w.pin_sam_caps[fl])) { w.pin_sam_caps[fl] = 0; return fail("pinned host allocation failed"); }
      }
      w.pin_sam = w.pin_sams[fl];
      b2_d2h(w.pin_sam, w.sam_out.p, (size_t)total, w.stream);
      w.d2h_bytes += (size_t)total;
      b2_g_d2h_bytes += (size_t)total;
      ev_d2h = w.timer.mark();
      if (b2_sync(w.stream)) return fail("device synchronisation failed");
      if (sam) sam->append((const char*)w.pin_sam, (size_t)total);
      w.last_sam_bytes = (size_t)total;
    } else {
      if (b2_sync(w.stream)) return fail("device synchronisation failed");
    }
    if (check_cuda("pipeline")) return 1;
    b2_g_launches += (unsigned long long)w.times.launches;
    w.times.seed = w.timer.ms(ev_begin, ev_seed);
    w.times.sa = w.timer.ms(ev_seed, ev_sa);
    w.times.chain = w.timer.ms(ev_sa, ev_chain);
    w.times.extend = w.timer.ms(ev_chain, ev_ext);
    w.times.pestat = w.timer.ms(ev_ext, ev_pes0);
    w.times.rescue = w.timer.ms(ev_pes0, ev_pes);
    w.times.final_ = w.timer.ms(ev_pes, ev_fin);
    w.times.gather = w.timer.ms(ev_fin, ev_gat);
    w.times.d2h = w.timer.ms(ev_gat, ev_d2h);
    w.times.total = w.timer.ms(ev_begin, ev_d2h);
    if (ingest) w.times.h2d = w.timer.ms(ev_in0, ev_in1);
    if (ev_x[1] >= 0) w.times.k_xext = w.timer.ms(ev_x[0], ev_x[1]) + (ev_x[3] >= 0 ? w.timer.ms(ev_x[2], ev_x[3]) : 0.f);
    if (ev_r[1] >= 0) w.times.k_resc_sw = w.timer.ms(ev_r[0], ev_r[1]);
    if (ev_g[1] >= 0) w.times.k_galn = w.timer.ms(ev_g[0], ev_g[1]);
    {  // publish what grew during this batch
      auto raise = [](auto& shared, auto v) { auto cur = shared.load(); while (cur < v && !shared.compare_exchange_weak(cur, v)) {} };
      if (w.scratch_per_lane > scratch_in) raise(grown_scratch, w.scratch_per_lane);
      if (w.big_slot > big_slot_in) raise(grown_big_slot, w.big_slot);
      if (w.cap_intv > cap_intv_in) raise(grown_cap_intv, w.cap_intv);
      if (w.slot_size > slot_in && slot_in) raise(grown_slot, w.slot_size);
    }
    if (overflowed) {
      if (w.tried_decay) w.decay_wait = std::min(w.decay_wait * 4, 4096);  // the smaller size was not enough: ask less often
      w.clean_batches = 0;
    } else {
      if (w.tried_decay) w.decay_wait = 4;
      ++w.clean_batches;
    }
    w.tried_decay = false;
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------
// Insert-size statistics of one batch (what mem_pestat computes, bwa/bwamem_pair.c:46-109), from the per-pair candidates
// k_pestat left: each orientation's sizes are reduced to ascending (value, multiplicity) runs; quartiles are read off the
// cumulative multiplicities, the mean from exact integer sums, and the variance by adding each run's squared deviation
// `multiplicity` times in ascending order -- the same sequence of double additions as a pass over the sorted list, which
// is what keeps avg/std (and the low/high bounds derived from them) bit-identical.
namespace {
struct B2IsRun { uint64_t v; size_t cnt; };
struct B2IsDist {
  std::vector<B2IsRun> runs;
  size_t total = 0;
  void build(std::vector<uint64_t>& vals) {
    std::sort(vals.begin(), vals.end());
    total = vals.size();
    for (size_t i = 0; i < vals.size();) {
      size_t j = i;
      while (j < vals.size() && vals[j] == vals[i]) ++j;
      runs.push_back(B2IsRun{vals[i], j - i});
      i = j;
    }
  }
  uint64_t at_rank(size_t k) const {  // k-th smallest (0-based)
    for (const B2IsRun& r : runs) { if (k < r.cnt) return r.v; k -= r.cnt; }
    return runs.empty() ? 0 : runs.back().v;
  }
  int quantile(double f) const { return (int)at_rank((size_t)(int)(f * total + .499)); }
};
const char* b2_orient_name(int d) { static const char* nm[4] = {"FF", "FR", "RF", "RR"}; return nm[d]; }
}  // namespace

void b2_pestat_host(const B2Opt& opt, int n_pairs, const int8_t* dir, const int64_t* is, B2Pes pes[4], int verbose) {
  (void)opt;
  const bool say = verbose >= 3;
  B2IsDist dist[4];
  {
    std::vector<uint64_t> vals[4];
    for (int i = 0; i < n_pairs; ++i)
      if (dir[i] >= 0) vals[dir[i]].push_back((uint64_t)is[i]);
    for (int d = 0; d < 4; ++d) dist[d].build(vals[d]);
  }
  if (say)
    fprintf(stderr, "[M::mem_pestat] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n",
            (long)dist[0].total, (long)dist[1].total, (long)dist[2].total, (long)dist[3].total);
  size_t most = 0;
  for (int d = 0; d < 4; ++d) {
    const B2IsDist& D = dist[d];
    B2Pes& out = pes[d];
    memset(&out, 0, sizeof(B2Pes));
    most = std::max(most, D.total);
    if (D.total < 10) {
      if (say) fprintf(stderr, "[M::mem_pestat] skip orientation %s as there are not enough pairs\n", b2_orient_name(d));
      out.failed = 1;
      continue;
    }
    if (say) fprintf(stderr, "[M::mem_pestat] analyzing insert size distribution for orientation %s...\n", b2_orient_name(d));
    const int q1 = D.quantile(.25), q2 = D.quantile(.50), q3 = D.quantile(.75), iqr = q3 - q1;
    // inliers for the moments: within 2 IQR of the quartiles
    int in_lo = (int)(q1 - 2.0 * iqr + .499), in_hi = (int)(q3 + 2.0 * iqr + .499);
    if (in_lo < 1) in_lo = 1;
    if (say) {
      fprintf(stderr, "[M::mem_pestat] (25, 50, 75) percentile: (%d, %d, %d)\n", q1, q2, q3);
      fprintf(stderr, "[M::mem_pestat] low and high boundaries for computing mean and std.dev: (%d, %d)\n", in_lo, in_hi);
    }
    double sum = 0;   // integers below 2^53: exact whatever the grouping
    size_t n_in = 0;
    for (const B2IsRun& r : D.runs)
      if (r.v >= (uint64_t)in_lo && r.v <= (uint64_t)in_hi) { sum += (double)r.v * (double)r.cnt; n_in += r.cnt; }
    const double mean = sum / (int)n_in;
    double ss = 0;
    for (const B2IsRun& r : D.runs)
      if (r.v >= (uint64_t)in_lo && r.v <= (uint64_t)in_hi) {
        const double dev2 = ((double)r.v - mean) * ((double)r.v - mean);
        for (size_t k = 0; k < r.cnt; ++k) ss += dev2;   // one addition per pair, in ascending order
      }
    const double sd = sqrt(ss / (int)n_in);
    if (say) fprintf(stderr, "[M::mem_pestat] mean and std.dev: (%.2f, %.2f)\n", mean, sd);
    // proper-pair window: 3 IQR around the quartiles, widened to mean +- 4 sd
    int lo = (int)(q1 - 3.0 * iqr + .499), hi = (int)(q3 + 3.0 * iqr + .499);
    if (lo > mean - 4.0 * sd) lo = (int)(mean - 4.0 * sd + .499);
    if (hi < mean + 4.0 * sd) hi = (int)(mean + 4.0 * sd + .499);
    if (lo < 1) lo = 1;
    if (say) fprintf(stderr, "[M::mem_pestat] low and high boundaries for proper pairs: (%d, %d)\n", lo, hi);
    out.low = lo; out.high = hi; out.avg = mean; out.std = sd;
  }
  for (int d = 0; d < 4; ++d)   // an orientation with under 5 % of the commonest one's pairs is not trusted
    if (!pes[d].failed && dist[d].total < most * 0.05) {
      pes[d].failed = 1;
      if (say) fprintf(stderr, "[M::mem_pestat] skip orientation %s\n", b2_orient_name(d));
    }
}

// ---------------------------------------------------------------------------------------------
int b2_device_count() {
#if defined(B2_HOSTSIM)
  const char* s = getenv("B200BWA_HOSTSIM_DEVICES");  // the host checker build pretends to have this many devices
  return s ? atoi(s) : 1;
#else
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
#endif
}

B2Engine::B2Engine() : impl_(new B2EngineImpl) {}
B2Engine::~B2Engine() { delete impl_; }
int B2Engine::init(const B2IndexHost& idx, const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg) {
  return impl_->init(idx, opt, rg_id, cfg);
}
int B2Engine::init_device_index(const B2IndexHost& meta, const void* d_bwt, const void* d_sa, const void* d_pac,
                                const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg) {
  return impl_->init_device(meta, d_bwt, d_sa, d_pac, opt, rg_id, cfg);
}
int B2Engine::init_from_peer(const B2IndexHost& meta, B2Engine& src, const B2Opt& opt, const char* rg_id, const B2EngineConfig& cfg) {
  return impl_->init_peer(meta, *src.impl_, opt, rg_id, cfg);
}
int B2Engine::device() const { return impl_->cfg.device; }
int B2Engine::process(const B2HostBatch& b, const B2Pes* pes0, std::string& sam, int worker) {
  Worker& w = impl_->workers[worker];
  if (impl_->upload(w, b)) return 1;
  return impl_->run(w, pes0, &sam);
}
int B2Engine::process_view(const B2HostBatch& b, const B2Pes* pes0, int worker, const char** data, size_t* len) {
  Worker& w = impl_->workers[worker];
  *data = 0; *len = 0;
  if (impl_->upload(w, b)) return 1;
  if (impl_->run(w, pes0, nullptr, nullptr, 0, -1, true)) return 1;
  *data = (const char*)w.pin_sam; *len = w.last_sam_bytes;
  return 0;
}
int B2Engine::process_raw(const B2RawInput& in, const B2Pes* pes0, std::string& sam, int worker) {
  Worker& w = impl_->workers[worker];
  if (impl_->upload_raw(w, in)) return 1;
  return impl_->run(w, pes0, &sam);
}
int B2Engine::process_raw_view(const B2RawInput& in, const B2Pes* pes0, int worker, const char** data, size_t* len) {
  Worker& w = impl_->workers[worker];
  *data = 0; *len = 0;
  if (impl_->upload_raw(w, in)) return 1;
  if (impl_->run(w, pes0, nullptr, nullptr, 0, -1, true)) return 1;
  *data = (const char*)w.pin_sam; *len = w.last_sam_bytes;
  return 0;
}
int B2Engine::next_view_slot(int worker) const { return impl_->workers[worker].sam_flip; }
int B2Engine::process_items(const B2HostBatch& b, const B2Pes* pes0, int worker, int paired, std::string& sam,
                            std::vector<int32_t>& item_len) {
  Worker& w = impl_->workers[worker];
  if (impl_->upload(w, b)) return 1;
  return impl_->run(w, pes0, &sam, nullptr, 0, -1, false, paired ? 1 : 0, &item_len);
}
int B2Engine::upload(const B2HostBatch& b, int worker) { return impl_->upload(impl_->workers[worker], b); }
int B2Engine::run_resident(const B2Pes* pes0, std::string* sam, int worker) {
  return impl_->run(impl_->workers[worker], pes0, sam);
}
int B2Engine::run_resident_range(const B2Pes* pes0, std::string* sam, int worker, int src_worker, int first, int count, int paired) {
  return impl_->run(impl_->workers[worker], pes0, sam, &impl_->workers[src_worker], first, count, false, paired < 0 ? -1 : (paired ? 1 : 0));
}
void B2Engine::set_options(const B2Opt& opt, const char* rg_id, int verbose) {
  impl_->bind_device();
  impl_->opt = opt; impl_->cfg.verbose = verbose; impl_->set_rg(rg_id);
}
const B2Opt& B2Engine::options() const { return impl_->opt; }
const B2StageTimes& B2Engine::times(int worker) const { return impl_->workers[worker].times; }
int B2Engine::counters(int worker, unsigned long long out[16], unsigned long long* h2d, unsigned long long* d2h, int reset) {
  impl_->bind_device();
  Worker& w = impl_->workers[worker];
  b2_d2h(out, w.counters.p, 16 * sizeof(unsigned long long), w.stream);
  b2_sync(w.stream);
  if (h2d) *h2d = w.h2d_bytes;
  if (d2h) *d2h = w.d2h_bytes;
  if (reset) {
    b2_dev_memset(w.counters.p, 0, 16 * sizeof(unsigned long long), w.stream);
    b2_sync(w.stream);
    w.h2d_bytes = w.d2h_bytes = 0;
  }
  return 0;
}
const char* B2Engine::last_error() const {
  static thread_local std::string copy;  // the engine's string may be rewritten by a failing worker thread
  copy = impl_->last_error();
  return copy.c_str();
}
int B2Engine::n_workers() const { return (int)impl_->workers.size(); }
void B2Engine::index_device_ptrs(void** bwt, size_t* bb, void** sa, size_t* sb, void** pac, size_t* pb) {
  *bwt = (void*)impl_->ix.bwt; *bb = impl_->bwt_bytes; *sa = (void*)impl_->ix.sa; *sb = impl_->sa_bytes;
  *pac = (void*)impl_->ix.pac; *pb = impl_->pac_bytes;
}
int B2Engine::dump_intervals(int worker, std::vector<int32_t>& n_intv, std::vector<B2Intv>& intv) {
  Worker& w = impl_->workers[worker];
  n_intv.resize(w.n_reads); intv.resize((size_t)w.n_reads * w.cap_intv);
  b2_d2h(n_intv.data(), w.n_intv.p, w.n_reads * sizeof(int32_t), w.stream);
  b2_d2h(intv.data(), w.intv.p, intv.size() * sizeof(B2Intv), w.stream);
  b2_sync(w.stream);
  return w.cap_intv;
}
int B2Engine::dump_regs(int worker, std::vector<int32_t>& n_regs, std::vector<int64_t>& reg_off, std::vector<B2Reg>& regs) {
  Worker& w = impl_->workers[worker];
  n_regs.resize(w.n_reads); reg_off.resize(w.n_reads + 1);
  b2_d2h(n_regs.data(), w.n_regs.p, w.n_reads * sizeof(int32_t), w.stream);
  b2_d2h(reg_off.data(), w.reg_off.p, (w.n_reads + 1) * sizeof(int64_t), w.stream);
  b2_sync(w.stream);
  regs.resize(reg_off[w.n_reads]);
  b2_d2h(regs.data(), w.regs.p, regs.size() * sizeof(B2Reg), w.stream);
  b2_sync(w.stream);
  return 0;
}
