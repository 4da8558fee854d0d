// Host-side driver interface (option block, index loader, FASTA/Q batching, main_mem twin).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "b2_engine.h"

#define B200BWA_VERSION "0.7.15-r1140-b200"

struct B2MemArgs {
  B2Opt opt;
  int verbose, copy_comment, fixed_chunk_size, ignore_alt, has_pes0;
  int no_mt_io;   // -1: one I/O thread in the reference's pipeline (decides how a malformed record ends the input, b2_mem_main)
  B2Pes pes0[4];
  std::string rg_id, hdr_line;
  std::vector<std::string> positional;  // idxbase, in1 [, in2]
};

void b2_opt_init(B2Opt* o);                                    // mem_opt_init, bwa/bwamem.c:48
void b2_fill_scmat(int a, int b, int8_t mat[25]);              // bwa_fill_scmat, bwa/bwa.c:99
int b2_parse_mem_args(int argc, char** argv, B2MemArgs* m);    // getopt loop + presets, bwa/fastmap.c:133-318
int b2_load_index(const char* hint, B2IndexHost* ix, int verbose, int meta_only);  // bwa/bwa.c:252
std::string b2_sam_header(const B2IndexHost& ix, const std::string& hdr_line, const std::string& pg);  // bwa/bwa.c:370
const unsigned char* b2_nt4_table();

// One FASTA/Q record as pointers (into a memory-mapped input file, or into storage owned by the batch).
struct B2RecView {
  const char *name, *comment, *seq, *qual;
  int l_name, l_comment, l_seq, l_qual;
};

// One -K batch as record views: what the reader thread produces; a GPU worker thread packs it (b2_pack_batch).
// Two forms: `compact` (every record is a plain 4-line FASTQ record of a mapped file: descriptors only, the device
// does the packing) or views (anything else: packed on the host by b2_pack_batch).
struct B2RawBatch {
  std::vector<B2RecView> recs;
  int64_t n_bases = 0, n_processed = 0;
  bool compact = false;
  int n_files = 0, max_len = 0;
  const unsigned char* cbase[2] = {0, 0};
  std::vector<B2Rec> crec[2];
  size_t n_reads() const { return compact ? crec[0].size() + crec[1].size() : recs.size(); }
  std::vector<std::unique_ptr<char[]>> arena;  // copies of the records that do not live in a mapped file
  size_t arena_used = 0, arena_cap = 0;
  char* hold(const char* p, size_t n);
  void clear() {
    recs.clear(); arena.clear(); arena_used = arena_cap = 0; n_bases = 0;
    compact = false; n_files = 0; max_len = 0; crec[0].clear(); crec[1].clear(); cbase[0] = cbase[1] = 0;
  }
};

class B2SeqReader {  // kseq_read semantics, bwa/kseq.h:175-215
 public:
  B2SeqReader();
  ~B2SeqReader();
  int open(const char* fn);
  void close();
  int next();  // >= 0 length, -1 EOF, -2 truncated quality
  // same, returning views; *stable: the views stay valid until close() (mapped file), else until the next call
  int next_view(B2RecView* v, bool* stable);
  // the records the scan-ahead has ready, as descriptors into the mapping (false: none, use next_view); advance(k)
  // consumes the first k of them
  bool span(const B2Rec** p, size_t* n);
  void advance(size_t k);
  const unsigned char* map_base() const;
  static void view_of(const unsigned char* base, const B2Rec& r, B2RecView* v);
  const std::string& name() const;
  const std::string& comment() const;
  const std::string& seq() const;
  const std::string& qual() const;
  // "" or the message of a failed gzread (a corrupt or cut-off .gz, a directory): fatal in the reference (err_gzread,
  // bwa/utils.c:214-225); the reader reports end of input and the caller turns this into status 1
  const std::string& io_error() const;
 private:
  struct Impl;
  Impl* impl_;
};

int b2_read_batch(int chunk_size, B2SeqReader* ks, B2SeqReader* ks2, int copy_comment, B2HostBatch* b);  // bseq_read
int b2_read_raw_batch(int chunk_size, B2SeqReader* ks, B2SeqReader* ks2, B2RawBatch* rb);                  // its cut, views only
void b2_pack_batch(const B2RawBatch& rb, int copy_comment, B2HostBatch* b);
void b2_engine_config_from_env(B2EngineConfig* cfg);
B2Engine* b2_get_engine(const char* prefix, int device, int ignore_alt, const B2Opt& opt, const char* rg_id, int verbose,
                        B2IndexHost** host_out);
int b2_index_cache_evict();  // releases every cached device image no call is using; returns how many
// argv[0] = "mem"; out_path null = stdout; pg_cmdline = full @PG line or null
int b2_mem_main(int argc, char** argv, const char* out_path, const char* pg_cmdline);

int b2_index_main(int argc, char** argv);   // argv[0] = "index": `bwa index` twin, BWT/Occ/SA built on the GPU
int b2_reduce_sam(int n_inputs, const char* const* inputs, const char* out_path, int delete_inputs);
int b2_run_partitions(int n_opts, char** opts, const char* idxbase, int n_parts, const char* const* fq1, const char* const* fq2,
                      const char* out_dir, const char* app, int reduce);

extern thread_local std::string b2_tls_error;
