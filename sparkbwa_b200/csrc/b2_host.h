// Host-side driver interface (option block, index loader, FASTA/Q batching, main_mem twin).
#pragma once
#include <string>
#include <vector>
#include "b2_engine.h"

#define B200BWA_VERSION "0.7.15-r1140-b200"

struct B2MemArgs {
  B2Opt opt;
  int verbose, copy_comment, fixed_chunk_size, ignore_alt, has_pes0;
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

class B2SeqReader {  // kseq_read semantics, bwa/kseq.h:175-215
 public:
  B2SeqReader();
  ~B2SeqReader();
  int open(const char* fn);
  void close();
  int next();  // >= 0 length, -1 EOF, -2 truncated quality
  const std::string& name() const;
  const std::string& comment() const;
  const std::string& seq() const;
  const std::string& qual() const;
 private:
  struct Impl;
  Impl* impl_;
};

int b2_read_batch(int chunk_size, B2SeqReader* ks, B2SeqReader* ks2, int copy_comment, B2HostBatch* b);  // bseq_read
void b2_engine_config_from_env(B2EngineConfig* cfg);
B2Engine* b2_get_engine(const char* prefix, int device, int ignore_alt, const B2Opt& opt, const char* rg_id, int verbose,
                        B2IndexHost** host_out);
// argv[0] = "mem"; out_path null = stdout; pg_cmdline = full @PG line or null
int b2_mem_main(int argc, char** argv, const char* out_path, const char* pg_cmdline);

extern thread_local std::string b2_tls_error;
