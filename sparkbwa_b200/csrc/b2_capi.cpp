// C-ABI entry points (include/b200bwa.h).  b200bwa_main replaces bwa's main() for the `mem`
// sub-command (bwa/main.c:61-104): it builds the @PG line from argv (main.c:68-70) and dispatches
// to the main_mem twin; every other sub-command is reported as unrecognised with return code 1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <memory>
#include <atomic>
#include "../../include/b200bwa.h"
#include "b2_host.h"

static std::string make_pg(int argc, char** argv) {
  std::string pg = "@PG\tID:bwa\tPN:bwa\tVN:" B200BWA_VERSION "\tCL:";
  pg += argc > 0 ? argv[0] : "bwa";
  for (int i = 1; i < argc; ++i) { pg += ' '; pg += argv[i]; }
  return pg;
}

extern "C" {

const char* b200bwa_last_error(void) { return b2_tls_error.c_str(); }

int b200bwa_mem(int argc, char** argv, const char* out_path) {
  try {
    b2_tls_error.clear();
    std::vector<char*> full;
    std::string prog = "bwa";
    full.push_back(&prog[0]);
    for (int i = 0; i < argc; ++i) full.push_back(argv[i]);
    std::string pg = make_pg((int)full.size(), full.data());
    return b2_mem_main(argc, argv, out_path, pg.c_str());
  } catch (const std::exception& e) {
    b2_tls_error = e.what();
    fprintf(stderr, "[E::b200bwa] %s\n", e.what());
    return 1;
  }
}

int b200bwa_main_to(int argc, char** argv, const char* out_path) {
  try {
    b2_tls_error.clear();
    if (argc < 2) {
      fprintf(stderr, "\nProgram: b200bwa (B200-native BWA-MEM behind the SparkBWA bridge)\nVersion: %s\n\nUsage:   bwa mem [options] <idxbase> <in1.fq> [in2.fq]\n\n", B200BWA_VERSION);
      return 1;
    }
    if (strcmp(argv[1], "index") == 0) return b2_index_main(argc - 1, argv + 1);
    if (strcmp(argv[1], "mem") != 0) {
      fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
      b2_tls_error = std::string("unrecognized command '") + argv[1] + "' (`mem` and `index` are provided by b200bwa)";
      return 1;
    }
    std::string pg = make_pg(argc, argv);
    return b2_mem_main(argc - 1, argv + 1, out_path, pg.c_str());
  } catch (const std::exception& e) {
    b2_tls_error = e.what();
    fprintf(stderr, "[E::b200bwa] %s\n", e.what());
    return 1;
  }
}

int b200bwa_main(int argc, char** argv) { return b200bwa_main_to(argc, argv, 0); }

int b200bwa_index_cache_evict(void) { return b2_index_cache_evict(); }

int b200bwa_index_build(int argc, char** argv) {
  try { b2_tls_error.clear(); return b2_index_main(argc, argv); }
  catch (const std::exception& e) { b2_tls_error = e.what(); fprintf(stderr, "[E::b200bwa] %s\n", e.what()); return 1; }
}

int b200bwa_reduce_sam(int n_inputs, const char* const* inputs, const char* out_path, int delete_inputs) {
  try { b2_tls_error.clear(); return b2_reduce_sam(n_inputs, inputs, out_path, delete_inputs); }
  catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

int b200bwa_run_partitions(int n_opts, char** opts, const char* idxbase, int n_parts, const char* const* fq1,
                           const char* const* fq2, const char* out_dir, const char* app_name, int reduce) {
  try { b2_tls_error.clear(); return b2_run_partitions(n_opts, opts, idxbase, n_parts, fq1, fq2, out_dir, app_name, reduce); }
  catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Library-level API
struct b200bwa_index {
  std::string prefix;
  int device;
  B2IndexHost host;
  B2Engine* engine;
  B2MemArgs args;
  bool adopted;
};

static int split_opts(const char* s, std::vector<std::string>& tok) {
  tok.clear();
  tok.push_back("mem");
  if (!s) return 0;
  std::string cur;
  for (const char* p = s;; ++p) {
    if (*p == ' ' || *p == 0) {
      if (!cur.empty()) { tok.push_back(cur); cur.clear(); }
      if (!*p) break;
    } else cur += *p;
  }
  return 0;
}

static int fill_batch(B2HostBatch& b, int n_reads, const char* seq, const char* qual, const int64_t* seq_off,
                      const char* names, int64_t n_processed) {
  const unsigned char* t = b2_nt4_table();
  b.clear();
  b.n_reads = n_reads;
  b.n_processed = n_processed;
  const int64_t total = seq_off[n_reads];
  b.seq.resize(total); b.qual.resize(total);
  for (int64_t i = 0; i < total; ++i) b.seq[i] = t[(unsigned char)seq[i]];
  if (qual) memcpy(b.qual.data(), qual, total);
  b.has_qual.assign(n_reads, qual ? 1 : 0);
  b.seq_off.assign(seq_off, seq_off + n_reads);
  b.l_seq.resize(n_reads); b.name_off.resize(n_reads); b.l_name.resize(n_reads);
  b.com_off.assign(n_reads, 0); b.l_com.assign(n_reads, -1);
  const char* p = names;
  for (int i = 0; i < n_reads; ++i) {
    int l = (int)(seq_off[i + 1] - seq_off[i]);
    b.l_seq[i] = l;
    if (l > b.max_len) b.max_len = l;
    size_t ln = strlen(p);
    b.name_off[i] = (int64_t)b.names.size();
    b.l_name[i] = (int32_t)ln;
    b.names.insert(b.names.end(), p, p + ln);
    p += ln + 1;
  }
  b.n_bases = total;
  return 0;
}

static char* dup_out(const std::string& s, size_t* len) {
  char* p = (char*)malloc(s.size() + 1);
  if (!p) return 0;
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  if (len) *len = s.size();
  return p;
}

extern "C" {

b200bwa_index_t* b200bwa_index_load(const char* prefix, int device) {
  try {
    b2_tls_error.clear();
    std::unique_ptr<b200bwa_index> ix(new b200bwa_index);
    ix->prefix = prefix; ix->device = device; ix->engine = 0; ix->adopted = false;
    std::vector<std::string> tok;
    split_opts(0, tok);
    std::vector<char*> av;
    for (auto& s : tok) av.push_back(&s[0]);
    if (b2_parse_mem_args((int)av.size(), av.data(), &ix->args)) return 0;
    ix->args.verbose = 2;
    if (b2_load_index(prefix, &ix->host, ix->args.verbose, 0)) return 0;
    ix->engine = new B2Engine;
    B2EngineConfig cfg;
    cfg.device = device; cfg.verbose = ix->args.verbose;
    b2_engine_config_from_env(&cfg);
    cfg.device = device;
    if (ix->engine->init(ix->host, ix->args.opt, "", cfg)) {
      b2_tls_error = ix->engine->last_error();
      delete ix->engine;
      return 0;
    }
    std::vector<uint32_t>().swap(ix->host.bwt);
    std::vector<uint64_t>().swap(ix->host.sa);
    std::vector<uint8_t>().swap(ix->host.pac);
    return ix.release();
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 0; }
}

b200bwa_index_t* b200bwa_index_adopt(const char* prefix, int device, const void* d_bwt, const void* d_sa, const void* d_pac) {
  try {
    b2_tls_error.clear();
    std::unique_ptr<b200bwa_index> ix(new b200bwa_index);
    ix->prefix = prefix; ix->device = device; ix->engine = 0; ix->adopted = true;
    std::vector<std::string> tok;
    split_opts(0, tok);
    std::vector<char*> av;
    for (auto& s : tok) av.push_back(&s[0]);
    if (b2_parse_mem_args((int)av.size(), av.data(), &ix->args)) return 0;
    ix->args.verbose = 2;
    if (b2_load_index(prefix, &ix->host, ix->args.verbose, 1)) return 0;
    ix->engine = new B2Engine;
    B2EngineConfig cfg;
    cfg.device = device; cfg.verbose = ix->args.verbose;
    b2_engine_config_from_env(&cfg);
    cfg.device = device;
    if (ix->engine->init_device_index(ix->host, d_bwt, d_sa, d_pac, ix->args.opt, "", cfg)) {
      b2_tls_error = ix->engine->last_error();
      delete ix->engine;
      return 0;
    }
    return ix.release();
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 0; }
}

void b200bwa_index_free(b200bwa_index_t* ix) {
  if (!ix) return;
  delete ix->engine;
  delete ix;
}

int b200bwa_set_options(b200bwa_index_t* ix, const char* mem_options) {
  if (!ix) return 1;
  std::vector<std::string> tok;
  split_opts(mem_options, tok);
  std::vector<char*> av;
  for (auto& s : tok) av.push_back(&s[0]);
  B2MemArgs a;
  if (b2_parse_mem_args((int)av.size(), av.data(), &a)) { b2_tls_error = "invalid bwa-mem option string"; return 1; }
  if (!a.positional.empty()) { b2_tls_error = "unexpected positional argument in option string"; return 1; }
  ix->args = a;
  ix->engine->set_options(a.opt, a.rg_id.c_str(), a.verbose);
  return 0;
}

int b200bwa_align_batch(b200bwa_index_t* ix, int n_reads, const char* seq, const char* qual, const int64_t* seq_off,
                        const char* names, int64_t n_processed, int paired, char** sam_out, size_t* sam_len) {
  try {
    if (!ix || !sam_out) return 1;
    b2_tls_error.clear();
    B2HostBatch b;
    fill_batch(b, n_reads, seq, qual, seq_off, names, n_processed);
    B2Opt o = ix->args.opt;
    if (paired) o.flag |= B2_F_PE; else o.flag &= ~B2_F_PE;
    ix->engine->set_options(o, ix->args.rg_id.c_str(), ix->args.verbose);
    std::string sam;
    if (ix->engine->process(b, ix->args.has_pes0 ? ix->args.pes0 : 0, sam, 0)) { b2_tls_error = ix->engine->last_error(); return 1; }
    *sam_out = dup_out(sam, sam_len);
    return *sam_out ? 0 : 1;
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

void b200bwa_free(void* p) { free(p); }

int b200bwa_sam_header(b200bwa_index_t* ix, const char* cmdline, char** out, size_t* out_len) {
  if (!ix || !out) return 1;
  std::string pg;
  if (cmdline) { pg = "@PG\tID:bwa\tPN:bwa\tVN:" B200BWA_VERSION "\tCL:"; pg += cmdline; }
  std::string h = b2_sam_header(ix->host, ix->args.hdr_line, pg);
  *out = dup_out(h, out_len);
  return *out ? 0 : 1;
}

int b200bwa_last_times(b200bwa_index_t* ix, b200bwa_times_t* t) {
  if (!ix || !t) return 1;
  const B2StageTimes& s = ix->engine->times(0);
  t->h2d = s.h2d; t->seed = s.seed; t->sa = s.sa; t->chain = s.chain; t->extend = s.extend; t->pestat = s.pestat;
  t->finalize = s.final_; t->gather = s.gather; t->d2h = s.d2h; t->total = s.total; t->launches = s.launches;
  t->n_seed_tasks = s.n_seed_tasks; t->n_regs = s.n_regs; t->sam_bytes = s.sam_bytes; t->rescue = s.rescue; t->n_rescue_tasks = s.n_resc;
  t->k_xext = s.k_xext; t->k_resc_sw = s.k_resc_sw; t->k_galn = s.k_galn;
  return 0;
}

int b200bwa_upload_batch(b200bwa_index_t* ix, int n_reads, const char* seq, const char* qual, const int64_t* seq_off,
                         const char* names, int64_t n_processed) {
  try {
    if (!ix) return 1;
    B2HostBatch b;
    fill_batch(b, n_reads, seq, qual, seq_off, names, n_processed);
    if (ix->engine->upload(b, 0)) { b2_tls_error = ix->engine->last_error(); return 1; }
    return 0;
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

int b200bwa_run_resident(b200bwa_index_t* ix, int paired, int want_sam, char** sam_out, size_t* sam_len) {
  try {
    if (!ix) return 1;
    B2Opt o = ix->args.opt;
    if (paired) o.flag |= B2_F_PE; else o.flag &= ~B2_F_PE;
    ix->engine->set_options(o, ix->args.rg_id.c_str(), ix->args.verbose);
    std::string sam;
    if (ix->engine->run_resident(ix->args.has_pes0 ? ix->args.pes0 : 0, want_sam ? &sam : 0, 0)) {
      b2_tls_error = ix->engine->last_error();
      return 1;
    }
    if (want_sam && sam_out) { *sam_out = dup_out(sam, sam_len); return *sam_out ? 0 : 1; }
    return 0;
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

int b200bwa_run_resident_range(b200bwa_index_t* ix, int worker, int first, int count, int paired, int want_sam,
                               char** sam_out, size_t* sam_len, b200bwa_times_t* t) {
  try {
    if (!ix) return 1;
    if (worker < 0 || worker >= ix->engine->n_workers()) {
      b2_tls_error = "worker index out of range (the engine has " + std::to_string(ix->engine->n_workers()) + " stream/buffer sets)";
      return 1;
    }
    // options are set once by the caller (b200bwa_set_options / b200bwa_set_paired); concurrent workers share them
    std::string sam;
    if (ix->engine->run_resident_range(ix->args.has_pes0 ? ix->args.pes0 : 0, want_sam ? &sam : 0, worker, 0, first, count, paired ? 1 : 0)) {
      b2_tls_error = ix->engine->last_error();
      return 1;
    }
    if (t) {
      const B2StageTimes& s = ix->engine->times(worker);
      t->h2d = s.h2d; t->seed = s.seed; t->sa = s.sa; t->chain = s.chain; t->extend = s.extend; t->pestat = s.pestat;
      t->finalize = s.final_; t->gather = s.gather; t->d2h = s.d2h; t->total = s.total; t->launches = s.launches;
      t->n_seed_tasks = s.n_seed_tasks; t->n_regs = s.n_regs; t->sam_bytes = s.sam_bytes; t->rescue = s.rescue; t->n_rescue_tasks = s.n_resc;
      t->k_xext = s.k_xext; t->k_resc_sw = s.k_resc_sw; t->k_galn = s.k_galn;
    }
    if (want_sam && sam_out) { *sam_out = dup_out(sam, sam_len); return *sam_out ? 0 : 1; }
    return 0;
  } catch (const std::exception& e) { b2_tls_error = e.what(); return 1; }
}

int b200bwa_n_workers(b200bwa_index_t* ix) { return ix ? ix->engine->n_workers() : 0; }

int b200bwa_set_paired(b200bwa_index_t* ix, int paired) {
  if (!ix) return 1;
  if (paired) ix->args.opt.flag |= B2_F_PE; else ix->args.opt.flag &= ~B2_F_PE;
  ix->engine->set_options(ix->args.opt, ix->args.rg_id.c_str(), ix->args.verbose);
  return 0;
}

int b200bwa_counters(b200bwa_index_t* ix, int worker, unsigned long long out[16], int reset) {
  if (!ix || worker < 0 || worker >= ix->engine->n_workers()) return 1;
  return ix->engine->counters(worker, out, 0, 0, reset);
}

extern std::atomic<unsigned long long> b2_g_h2d_bytes, b2_g_d2h_bytes, b2_g_launches;
int b200bwa_process_counters(unsigned long long out[3], int reset) {
  out[0] = b2_g_h2d_bytes.load(); out[1] = b2_g_d2h_bytes.load(); out[2] = b2_g_launches.load();
  if (reset) { b2_g_h2d_bytes = 0; b2_g_d2h_bytes = 0; b2_g_launches = 0; }
  return 0;
}

int b200bwa_index_device_image(b200bwa_index_t* ix, void** bwt, size_t* bwt_bytes, void** sa, size_t* sa_bytes, void** pac,
                               size_t* pac_bytes) {
  if (!ix) return 1;
  ix->engine->index_device_ptrs(bwt, bwt_bytes, sa, sa_bytes, pac, pac_bytes);
  return 0;
}

}  // extern "C"
