// Host driver: option parsing, index loading, FASTA/Q batching, SAM header and the batch pipeline
// around the GPU engine.  Mirrors the behaviour (not the code) of bwa/fastmap.c:115-364 (main_mem),
// :38-97 (process), bwa/bwa.c:42-75 (bseq_read), :252-284 (index load), :370-450 (SAM header, -R/-H),
// bwa/bntseq.c:98-206 (bns_restore) and bwa/kseq.h:175-215 (kseq_read).
#include "b2_host.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <unistd.h>
#include <sys/time.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

thread_local std::string b2_tls_error;

static int set_err(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  b2_tls_error = buf;
  return 1;
}

static double b2_realtime() {
  struct timeval tp;
  gettimeofday(&tp, 0);
  return tp.tv_sec + tp.tv_usec * 1e-6;
}
static double b2_cputime() {
  struct rusage r;
  getrusage(RUSAGE_SELF, &r);
  return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

// ---------------------------------------------------------------------------------------------
// Options
void b2_opt_init(B2Opt* o) {
  memset(o, 0, sizeof(*o));
  o->a = 1; o->b = 4;
  o->o_del = o->o_ins = 6;
  o->e_del = o->e_ins = 1;
  o->w = 100; o->T = 30; o->zdrop = 100;
  o->pen_unpaired = 17;
  o->pen_clip5 = o->pen_clip3 = 5;
  o->max_mem_intv = 20;
  o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500;
  o->max_chain_gap = 10000; o->max_ins = 10000;
  o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f;
  o->split_factor = 1.5f;
  o->chunk_size = 10000000; o->n_threads = 1;
  o->max_XA_hits = 5; o->max_XA_hits_alt = 200;
  o->max_matesw = 50;
  o->mask_level_redun = 0.95f;
  o->min_chain_weight = 0;
  o->max_chain_extend = 1 << 30;
  o->mapQ_coef_len = 50;
  o->mapQ_coef_fac = (int)log(o->mapQ_coef_len);
  b2_fill_scmat(o->a, o->b, o->mat);
}

void b2_fill_scmat(int a, int b, int8_t mat[25]) {
  int i, j, k;
  for (i = k = 0; i < 4; ++i) {
    for (j = 0; j < 4; ++j) mat[k++] = i == j ? a : -b;
    mat[k++] = -1;
  }
  for (j = 0; j < 5; ++j) mat[k++] = -1;
}

static void unescape(std::string& s) {  // bwa_escape: \t \n \r \\ (unknown escapes vanish)
  std::string o;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '\\') {
      ++i;
      if (i >= s.size()) break;
      if (s[i] == 't') o += '\t';
      else if (s[i] == 'n') o += '\n';
      else if (s[i] == 'r') o += '\r';
      else if (s[i] == '\\') o += '\\';
    } else o += s[i];
  }
  s.swap(o);
}

static void insert_header(const std::string& s, std::string& hdr) {
  if (s.empty() || s[0] != '@') return;
  std::string t = s;
  unescape(t);
  if (!hdr.empty()) hdr += '\n';
  hdr += t;
}

// Re-entrant scanner over the bwa-mem option string; options may follow positionals (glibc permutes).
int b2_parse_mem_args(int argc, char** argv, B2MemArgs* m) {
  static const char* spec = "1paMCSPVYjk:c:v:s:r:t:R:A:B:O:E:U:w:L:d:T:Q:D:m:I:N:W:x:G:h:y:K:X:H:";
  B2Opt* opt = &m->opt;
  B2Opt opt0;
  b2_opt_init(opt);
  memset(&opt0, 0, sizeof(opt0));
  m->verbose = 3; m->copy_comment = 0; m->fixed_chunk_size = -1; m->ignore_alt = 0; m->has_pes0 = 0; m->no_mt_io = 0;
  m->rg_id.clear(); m->hdr_line.clear(); m->positional.clear();
  memset(m->pes0, 0, sizeof(m->pes0));
  for (int i = 0; i < 4; ++i) m->pes0[i].failed = 1;
  std::string mode, rg_line;
  char* p;
  bool opts_done = false;
  for (int ai = 1; ai < argc; ++ai) {
    const char* arg = argv[ai];
    if (opts_done || arg[0] != '-' || arg[1] == 0) { m->positional.push_back(arg); continue; }
    if (arg[1] == '-' && arg[2] == 0) { opts_done = true; continue; }
    for (int ci = 1; arg[ci]; ++ci) {
      const int c = arg[ci];
      const char* sp = strchr(spec, c);
      if (!sp || c == ':') { fprintf(stderr, "%s: invalid option -- '%c'\n", argv[0], c); return 1; }
      const char* optarg = 0;
      if (sp[1] == ':') {
        if (arg[ci + 1]) optarg = arg + ci + 1;
        else if (ai + 1 < argc) optarg = argv[++ai];
        else { fprintf(stderr, "%s: option requires an argument -- '%c'\n", argv[0], c); return 1; }
      }
      if (c == 'k') opt->min_seed_len = atoi(optarg), opt0.min_seed_len = 1;
      else if (c == '1') m->no_mt_io = 1;
      else if (c == 'x') mode = optarg;
      else if (c == 'w') opt->w = atoi(optarg), opt0.w = 1;
      else if (c == 'A') opt->a = atoi(optarg), opt0.a = 1;
      else if (c == 'B') opt->b = atoi(optarg), opt0.b = 1;
      else if (c == 'T') opt->T = atoi(optarg), opt0.T = 1;
      else if (c == 'U') opt->pen_unpaired = atoi(optarg), opt0.pen_unpaired = 1;
      else if (c == 't') opt->n_threads = atoi(optarg), opt->n_threads = opt->n_threads > 1 ? opt->n_threads : 1;
      else if (c == 'P') opt->flag |= B2_F_NOPAIRING;
      else if (c == 'a') opt->flag |= B2_F_ALL;
      else if (c == 'p') opt->flag |= B2_F_PE | B2_F_SMARTPE;
      else if (c == 'M') opt->flag |= B2_F_NO_MULTI;
      else if (c == 'S') opt->flag |= B2_F_NO_RESCUE;
      else if (c == 'Y') opt->flag |= B2_F_SOFTCLIP;
      else if (c == 'V') opt->flag |= B2_F_REF_HDR;
      else if (c == 'c') opt->max_occ = atoi(optarg), opt0.max_occ = 1;
      else if (c == 'd') opt->zdrop = atoi(optarg), opt0.zdrop = 1;
      else if (c == 'v') m->verbose = atoi(optarg);
      else if (c == 'j') m->ignore_alt = 1;
      else if (c == 'r') opt->split_factor = atof(optarg), opt0.split_factor = 1.;
      else if (c == 'D') opt->drop_ratio = atof(optarg), opt0.drop_ratio = 1.;
      else if (c == 'm') opt->max_matesw = atoi(optarg), opt0.max_matesw = 1;
      else if (c == 's') opt->split_width = atoi(optarg), opt0.split_width = 1;
      else if (c == 'G') opt->max_chain_gap = atoi(optarg), opt0.max_chain_gap = 1;
      else if (c == 'N') opt->max_chain_extend = atoi(optarg), opt0.max_chain_extend = 1;
      else if (c == 'W') opt->min_chain_weight = atoi(optarg), opt0.min_chain_weight = 1;
      else if (c == 'y') opt->max_mem_intv = atol(optarg), opt0.max_mem_intv = 1;
      else if (c == 'C') m->copy_comment = 1;
      else if (c == 'K') m->fixed_chunk_size = atoi(optarg);
      else if (c == 'X') opt->mask_level = atof(optarg);
      else if (c == 'h') {
        opt0.max_XA_hits = opt0.max_XA_hits_alt = 1;
        opt->max_XA_hits = opt->max_XA_hits_alt = strtol(optarg, &p, 10);
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) opt->max_XA_hits_alt = strtol(p + 1, &p, 10);
      } else if (c == 'Q') {
        opt0.mapQ_coef_len = 1;
        opt->mapQ_coef_len = atoi(optarg);
        opt->mapQ_coef_fac = opt->mapQ_coef_len > 0 ? log(opt->mapQ_coef_len) : 0;
      } else if (c == 'O') {
        opt0.o_del = opt0.o_ins = 1;
        opt->o_del = opt->o_ins = strtol(optarg, &p, 10);
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) opt->o_ins = strtol(p + 1, &p, 10);
      } else if (c == 'E') {
        opt0.e_del = opt0.e_ins = 1;
        opt->e_del = opt->e_ins = strtol(optarg, &p, 10);
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) opt->e_ins = strtol(p + 1, &p, 10);
      } else if (c == 'L') {
        opt0.pen_clip5 = opt0.pen_clip3 = 1;
        opt->pen_clip5 = opt->pen_clip3 = strtol(optarg, &p, 10);
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) opt->pen_clip3 = strtol(p + 1, &p, 10);
      } else if (c == 'R') {
        std::string s = optarg;
        if (s.compare(0, 3, "@RG") != 0) {
          if (m->verbose >= 1) fprintf(stderr, "[E::bwa_set_rg] the read group line is not started with @RG\n");
          return 1;
        }
        unescape(s);
        size_t pid = s.find("\tID:");
        if (pid == std::string::npos) {
          if (m->verbose >= 1) fprintf(stderr, "[E::bwa_set_rg] no ID at the read group line\n");
          return 1;
        }
        pid += 4;
        size_t q = pid;
        while (q < s.size() && s[q] != '\t' && s[q] != '\n') ++q;
        if (q - pid + 1 > 256) {
          if (m->verbose >= 1) fprintf(stderr, "[E::bwa_set_rg] @RG:ID is longer than 255 characters\n");
          return 1;
        }
        m->rg_id = s.substr(pid, q - pid);
        rg_line = s;
      } else if (c == 'H') {
        if (optarg[0] != '@') {
          FILE* fp = fopen(optarg, "r");
          if (fp) {
            std::vector<char> buf(0x10000);
            while (fgets(buf.data(), 0xffff, fp)) {
              size_t l = strlen(buf.data());
              if (l && buf[l - 1] == '\n') buf[l - 1] = 0;
              insert_header(buf.data(), m->hdr_line);
            }
            fclose(fp);
          }
        } else insert_header(optarg, m->hdr_line);
      } else if (c == 'I') {
        B2Pes* pes = m->pes0;
        m->has_pes0 = 1;
        pes[1].failed = 0;
        pes[1].avg = strtod(optarg, &p);
        pes[1].std = pes[1].avg * .1;
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) pes[1].std = strtod(p + 1, &p);
        pes[1].high = (int)(pes[1].avg + 4. * pes[1].std + .499);
        pes[1].low = (int)(pes[1].avg - 4. * pes[1].std + .499);
        if (pes[1].low < 1) pes[1].low = 1;
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) pes[1].high = (int)(strtod(p + 1, &p) + .499);
        if (*p != 0 && ispunct(*p) && isdigit(p[1])) pes[1].low = (int)(strtod(p + 1, &p) + .499);
        if (m->verbose >= 3)
          fprintf(stderr, "[M::main_mem] mean insert size: %.3f, stddev: %.3f, max: %d, min: %d\n", pes[1].avg,
                  pes[1].std, pes[1].high, pes[1].low);
      }
      if (sp[1] == ':') break;  // the rest of this token was the argument
    }
  }
  if (!rg_line.empty()) {  // already unescaped
    if (!m->hdr_line.empty()) m->hdr_line += '\n';
    m->hdr_line += rg_line;
  }
  if (opt->n_threads < 1) opt->n_threads = 1;
  if (!mode.empty()) {
    if (mode == "intractg") {
      if (!opt0.o_del) opt->o_del = 16;
      if (!opt0.o_ins) opt->o_ins = 16;
      if (!opt0.b) opt->b = 9;
      if (!opt0.pen_clip5) opt->pen_clip5 = 5;
      if (!opt0.pen_clip3) opt->pen_clip3 = 5;
    } else if (mode == "pacbio" || mode == "pbref" || mode == "ont2d") {
      if (!opt0.o_del) opt->o_del = 1;
      if (!opt0.e_del) opt->e_del = 1;
      if (!opt0.o_ins) opt->o_ins = 1;
      if (!opt0.e_ins) opt->e_ins = 1;
      if (!opt0.b) opt->b = 1;
      if (opt0.split_factor == 0.) opt->split_factor = 10.;
      if (mode == "ont2d") {
        if (!opt0.min_chain_weight) opt->min_chain_weight = 20;
        if (!opt0.min_seed_len) opt->min_seed_len = 14;
      } else {
        if (!opt0.min_chain_weight) opt->min_chain_weight = 40;
        if (!opt0.min_seed_len) opt->min_seed_len = 17;
      }
      if (!opt0.pen_clip5) opt->pen_clip5 = 0;
      if (!opt0.pen_clip3) opt->pen_clip3 = 0;
    } else {
      fprintf(stderr, "[E::main_mem] unknown read type '%s'\n", mode.c_str());
      return 1;
    }
  } else if (opt0.a) {  // update_a: -A rescales the penalties the user did not set
    if (!opt0.b) opt->b *= opt->a;
    if (!opt0.T) opt->T *= opt->a;
    if (!opt0.o_del) opt->o_del *= opt->a;
    if (!opt0.e_del) opt->e_del *= opt->a;
    if (!opt0.o_ins) opt->o_ins *= opt->a;
    if (!opt0.e_ins) opt->e_ins *= opt->a;
    if (!opt0.zdrop) opt->zdrop *= opt->a;
    if (!opt0.pen_clip5) opt->pen_clip5 *= opt->a;
    if (!opt0.pen_clip3) opt->pen_clip3 *= opt->a;
    if (!opt0.pen_unpaired) opt->pen_unpaired *= opt->a;
  }
  // Values outside the options' domain.  With a match score of zero or less the reference dies in its first local
  // alignment (an assertion / a division by zero: bwa/ksw.c:348, :218); inside a JVM that must be a status, not a signal.
  // A negative -m reads as "no mate rescue" there (the loop bound of bwa/bwamem_pair.c:267 is never met): the same as 0.
  if (opt->a <= 0) {
    fprintf(stderr, "[E::main_mem] the matching score (-A) must be positive\n");
    return 1;
  }
  if (opt->max_matesw < 0) opt->max_matesw = 0;
  // Negative penalties, band width or seed length and a non-positive -c: the reference either dies (-w -1: abort, -c 0:
  // division by zero) or runs on with values its loops were not written for (-c -1 keeps no seed at all, -B -3 rewards
  // mismatches).  The kernels' packed cells and counts assume the documented domain, so these are refused up front.
  if (opt->b < 0 || opt->o_del < 0 || opt->o_ins < 0 || opt->e_del < 0 || opt->e_ins < 0 || opt->w < 0 || opt->min_seed_len < 0 ||
      opt->max_occ <= 0) {
    fprintf(stderr, "[E::main_mem] option value outside its domain (-B -O -E -w -k must not be negative, -c must be positive)\n");
    return 1;
  }
  b2_fill_scmat(opt->a, opt->b, opt->mat);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Index files
static int read_file(const std::string& fn, std::vector<char>& buf) {
  FILE* fp = fopen(fn.c_str(), "rb");
  if (!fp) return 1;
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  buf.resize(n);
  size_t got = n ? fread(buf.data(), 1, n, fp) : 0;
  fclose(fp);
  return got != (size_t)n;
}

static std::string infer_prefix(const std::string& hint) {
  FILE* fp;
  if ((fp = fopen((hint + ".64.bwt").c_str(), "rb")) != 0) { fclose(fp); return hint + ".64"; }
  if ((fp = fopen((hint + ".bwt").c_str(), "rb")) != 0) { fclose(fp); return hint; }
  return std::string();
}

int b2_load_index(const char* hint, B2IndexHost* ix, int verbose, int meta_only) {
  std::string prefix = infer_prefix(hint);
  if (prefix.empty()) {
    if (verbose >= 1) fprintf(stderr, "[E::bwa_idx_load_from_disk] fail to locate the index files\n");
    return set_err("fail to locate the index files for '%s'", hint);
  }
  {  // .bwt
    FILE* fp = fopen((prefix + ".bwt").c_str(), "rb");
    if (!fp) return set_err("cannot open %s.bwt", prefix.c_str());
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint64_t hdr[5];
    if (sz < 40 || fread(hdr, 8, 5, fp) != 5) { fclose(fp); return set_err("truncated %s.bwt", prefix.c_str()); }
    ix->primary = hdr[0];
    ix->L2[0] = 0;
    for (int i = 0; i < 4; ++i) ix->L2[i + 1] = hdr[i + 1];
    ix->seq_len = ix->L2[4];
    size_t words = (size_t)(sz - 40) >> 2;
    {  // a .bwt cut short (a partial copy to the executor's disk) would send the rank queries past the image: the file
       // must hold the 2-bit string plus one 32-byte count record per 128 symbols (bwt_bwtupdate_core, bwa/bwtindex.c:151-173)
      const uint64_t L = ix->seq_len;
      const uint64_t need = ((L + 15) >> 4) + ((L + 127) / 128 + 1) * 8;
      if (L == 0 || ix->primary > L || ix->L2[1] > ix->L2[2] || ix->L2[2] > ix->L2[3] || ix->L2[3] > L || (uint64_t)words < need) {
        fclose(fp);
        return set_err("truncated or corrupt %s.bwt (%zu words for %llu symbols)", prefix.c_str(), words, (unsigned long long)L);
      }
    }
    ix->bwt.resize(meta_only ? 0 : words);
    if (!meta_only && fread(ix->bwt.data(), 4, words, fp) != words) { fclose(fp); return set_err("truncated %s.bwt", prefix.c_str()); }
    ix->bwt_words = words;
    fclose(fp);
  }
  {  // .sa
    FILE* fp = fopen((prefix + ".sa").c_str(), "rb");
    if (!fp) return set_err("cannot open %s.sa", prefix.c_str());
    uint64_t hdr[7];
    if (fread(hdr, 8, 7, fp) != 7) { fclose(fp); return set_err("truncated %s.sa", prefix.c_str()); }
    if (hdr[0] != ix->primary) { fclose(fp); return set_err("SA-BWT inconsistency: primary is not the same."); }
    ix->sa_intv = (int)hdr[5];
    if (hdr[5] == 0 || hdr[5] > (1u << 20) || (hdr[5] & (hdr[5] - 1))) { fclose(fp); return set_err("corrupt %s.sa (sampling interval %llu)", prefix.c_str(), (unsigned long long)hdr[5]); }
    if (hdr[6] != ix->seq_len) { fclose(fp); return set_err("SA-BWT inconsistency: seq_len is not the same."); }
    size_t n_sa = (ix->seq_len + ix->sa_intv) / ix->sa_intv;
    ix->n_sa = n_sa;
    ix->sa.resize(meta_only ? 0 : n_sa);
    if (!meta_only) {
      ix->sa[0] = (uint64_t)-1;
      if (fread(ix->sa.data() + 1, 8, n_sa - 1, fp) != n_sa - 1) { fclose(fp); return set_err("truncated %s.sa", prefix.c_str()); }
    }
    fclose(fp);
  }
  {  // .ann
    FILE* fp = fopen((prefix + ".ann").c_str(), "r");
    if (!fp) return set_err("cannot open %s.ann", prefix.c_str());
    long long xx;
    unsigned seed;
    char str[8192];
    if (fscanf(fp, "%lld%d%u", &xx, &ix->n_seqs, &seed) != 3) { fclose(fp); return set_err("Parse error reading %s.ann", prefix.c_str()); }
    ix->l_pac = xx;
    ix->anns.resize(ix->n_seqs);
    ix->ann_names.resize(ix->n_seqs);
    ix->names.clear();
    for (int i = 0; i < ix->n_seqs; ++i) {
      B2Ann& p = ix->anns[i];
      unsigned gi;
      char* q = str;
      int c;
      if (fscanf(fp, "%u%8191s", &gi, str) != 2) { fclose(fp); return set_err("Parse error reading %s.ann", prefix.c_str()); }
      ix->ann_names[i] = str;
      p.name_off = (int32_t)ix->names.size(); p.name_len = (int32_t)strlen(str);
      ix->names += str;
      while (q - str < (long)sizeof(str) - 1 && (c = fgetc(fp)) != '\n' && c != EOF) *q++ = c;
      while (c != '\n' && c != EOF) c = fgetc(fp);
      if (c == EOF) { fclose(fp); return set_err("Unexpected end of file reading %s.ann", prefix.c_str()); }
      *q = 0;
      std::string anno;
      if (q - str > 1 && strcmp(str, " (null)") != 0) anno = str + 1;
      p.anno_off = (int32_t)ix->names.size(); p.anno_len = (int32_t)anno.size();
      ix->names += anno;
      int n_ambs;
      if (fscanf(fp, "%lld%d%d", &xx, &p.len, &n_ambs) != 3) { fclose(fp); return set_err("Parse error reading %s.ann", prefix.c_str()); }
      p.offset = xx;
      p.is_alt = 0;
    }
    fclose(fp);
  }
  {  // .amb: only consistency-checked (holes are not needed on the mem path)
    FILE* fp = fopen((prefix + ".amb").c_str(), "r");
    if (!fp) return set_err("cannot open %s.amb", prefix.c_str());
    long long xx;
    int n_seqs;
    if (fscanf(fp, "%lld%d%d", &xx, &n_seqs, &ix->n_holes) != 3) { fclose(fp); return set_err("Parse error reading %s.amb", prefix.c_str()); }
    fclose(fp);
    if (xx != ix->l_pac || n_seqs != ix->n_seqs) return set_err("inconsistent .ann and .amb files.");
    if ((uint64_t)ix->l_pac * 2 != ix->seq_len) return set_err("inconsistent .ann and .bwt files (%lld bases packed, %llu symbols indexed)", (long long)ix->l_pac, (unsigned long long)ix->seq_len);
    for (int i = 0; i < ix->n_seqs; ++i)   // contigs must tile [0, l_pac) (bns_pos2rid's binary search relies on it)
      if (ix->anns[i].offset < 0 || ix->anns[i].len < 0 || ix->anns[i].offset + ix->anns[i].len > ix->l_pac ||
          (i > 0 && ix->anns[i].offset < ix->anns[i - 1].offset + ix->anns[i - 1].len))
        return set_err("corrupt %s.ann (contig %d out of range)", prefix.c_str(), i);
  }
  if (!meta_only) {  // .pac
    FILE* fp = fopen((prefix + ".pac").c_str(), "rb");
    if (!fp) return set_err("cannot open %s.pac", prefix.c_str());
    ix->pac.resize(ix->l_pac / 4 + 1);
    if (fread(ix->pac.data(), 1, ix->pac.size(), fp) != ix->pac.size()) { fclose(fp); return set_err("truncated %s.pac", prefix.c_str()); }
    fclose(fp);
  }
  {  // .alt (optional)
    FILE* fp = fopen((prefix + ".alt").c_str(), "r");
    if (fp) {
      std::map<std::string, int> h;
      for (int i = 0; i < ix->n_seqs; ++i) h.insert(std::make_pair(ix->ann_names[i], i));  // first name wins (kh_put keeps the old key)
      // note: the reference overwrites the value on duplicates (kh_val(h,k) = i): last index wins
      for (int i = 0; i < ix->n_seqs; ++i) h[ix->ann_names[i]] = i;
      std::string str;
      int c;
      while ((c = fgetc(fp)) != EOF) {
        if (c == '\t' || c == '\n' || c == '\r') {
          if (!str.empty() && str[0] != '@') {
            std::map<std::string, int>::iterator it = h.find(str);
            if (it != h.end()) ix->anns[it->second].is_alt = 1;
          }
          while (c != '\n' && c != EOF) c = fgetc(fp);
          str.clear();
        } else str += (char)c;
      }
      fclose(fp);
    }
  }
  int c = 0;
  for (int i = 0; i < ix->n_seqs; ++i) if (ix->anns[i].is_alt) ++c;
  if (verbose >= 3) fprintf(stderr, "[M::bwa_idx_load_from_disk] read %d ALT contigs\n", c);
  return 0;
}

std::string b2_sam_header(const B2IndexHost& ix, const std::string& hdr_line, const std::string& pg) {
  std::string out;
  int n_SQ = 0;
  if (!hdr_line.empty()) {
    size_t p = 0;
    while ((p = hdr_line.find("@SQ\t", p)) != std::string::npos) {
      if (p == 0 || hdr_line[p - 1] == '\n') ++n_SQ;
      p += 4;
    }
  }
  if (n_SQ == 0) {
    char buf[64];
    for (int i = 0; i < ix.n_seqs; ++i) {
      out += "@SQ\tSN:"; out += ix.ann_names[i];
      snprintf(buf, sizeof(buf), "\tLN:%d", ix.anns[i].len);
      out += buf;
      out += ix.anns[i].is_alt ? "\tAH:*\n" : "\n";
    }
  } else if (n_SQ != ix.n_seqs)
    fprintf(stderr, "[W::bwa_print_sam_hdr] %d @SQ lines provided with -H; %d sequences in the index. Continue anyway.\n", n_SQ, ix.n_seqs);
  if (!hdr_line.empty()) { out += hdr_line; out += '\n'; }
  if (!pg.empty()) { out += pg; out += '\n'; }
  return out;
}

// ---------------------------------------------------------------------------------------------
// FASTA/FASTQ reader with kseq semantics
static const unsigned char nt4_table[256] = {
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};

const unsigned char* b2_nt4_table() { return nt4_table; }

// A record either lives in the reader's own strings (general kseq path: gzip, pipes, FASTA, multi-line
// records, CR/LF, ...) or, for the common case of a plain 4-line FASTQ in a regular file, is a set of
// pointers into the memory-mapped file found with four memchr calls (`stable`: valid until close()).
// The mapped fast path is taken record by record and only when the record has exactly the shape for
// which kseq_read's result is the four line bodies; anything else falls through to the general parser,
// which continues from the same position in the same buffer.
struct B2SeqReader::Impl {
  gzFile fp = 0;
  std::vector<unsigned char> store;      // gz mode: refill buffer
  const unsigned char* data = 0;          // current buffer (store.data() or the mapping)
  size_t begin = 0, end = 0;
  bool is_eof = false;
  std::string io_err;
  int last_char = 0;
  const unsigned char* map = 0;
  size_t map_size = 0;
  std::string name, comment, seq, qual;
  // Parallel scan-ahead (mapped files only).  The mapping is cut into fixed-size chunks; scanner threads take the
  // chunks in order, each finding the first record start inside its chunk speculatively (a line that parses as a plain
  // 4-line record and is followed by another one) and then parsing record after record into compact descriptors.  The
  // consumer accepts chunk k only if its start equals the position where chunk k-1 ended -- the scan from a given
  // position is deterministic, so an accepted chunk holds exactly the records a sequential scan would have found;
  // otherwise it re-scans the chunk itself from the right position.  The scan ends at EOF or at the first record that
  // is not a plain 4-line FASTQ record; the consumer then continues from that position with the synchronous code
  // (fast() / read() record by record).
  struct Chunk {
    size_t lo = 0, hi = 0, start = 0, end_pos = 0;
    bool stopped = false, done = false;   // stopped: an irregular record begins at end_pos
    std::vector<B2Rec> recs;
  };
  static constexpr size_t kNoStart = (size_t)-1;
  std::vector<std::thread> pre_threads;
  std::mutex pre_mu;
  std::condition_variable pre_cv;
  std::vector<std::unique_ptr<Chunk>> chunks;
  size_t chunk_bytes = (size_t)8 << 20, pre_next_chunk = 0, pre_consumed = 0, pre_window = 8;
  bool pre_on = false, pre_stop = false;
  size_t cur_chunk = 0, cur_i = 0, cur_expected = 0;
  bool cur_ready = false;

  // one plain 4-line record at `pos` (the conditions under which kseq_read's result is the four line bodies)
  static bool parse_rec(const unsigned char* data, size_t size, size_t pos, B2Rec* r, size_t* next) {
    const unsigned char* p = data + pos;
    const unsigned char* e = data + size;
    if (p >= e || *p != '@') return false;
    const unsigned char* nl = (const unsigned char*)memchr(p + 1, '\n', e - (p + 1));
    if (!nl || nl[-1] == '\r') return false;
    const unsigned char* q = p + 1;
    while (q < nl && !isspace(*q)) ++q;
    if (memchr(p + 1, 0, q - (p + 1))) return false;
    const unsigned char* s = nl + 1;
    if (s >= e || *s == '\n' || *s == '>' || *s == '+' || *s == '@') return false;
    const unsigned char* nl2 = (const unsigned char*)memchr(s, '\n', e - s);
    if (!nl2 || nl2[-1] == '\r') return false;
    const size_t ls = nl2 - s;
    if (memchr(s, 0, ls)) return false;
    const unsigned char* p3 = nl2 + 1;
    if (p3 >= e || *p3 != '+') return false;
    const unsigned char* nl3 = (const unsigned char*)memchr(p3, '\n', e - p3);
    if (!nl3) return false;
    const unsigned char* qs = nl3 + 1;
    if (qs + ls > e || (qs + ls < e && qs[ls] != '\n') || memchr(qs, '\n', ls)) return false;
    if ((size_t)(qs - p) > 0xffffffffu || ls > 0x7fffffffu) return false;
    r->off = pos; r->l_seq = (uint32_t)ls; r->l_hdr = (uint32_t)(nl - (p + 1)); r->l_name = (uint32_t)(q - (p + 1));
    r->qual_rel = (uint32_t)(qs - p);
    *next = (size_t)(qs + ls - data) + (qs + ls < e ? 1 : 0);
    return true;
  }
  static void rec_view(const unsigned char* data, const B2Rec& r, B2RecView* v) {
    const char* p = (const char*)data + r.off;
    v->name = p + 1; v->l_name = (int)r.l_name;
    if (r.l_name < r.l_hdr) { v->comment = p + 1 + r.l_name + 1; v->l_comment = (int)(r.l_hdr - r.l_name - 1); }
    else { v->comment = 0; v->l_comment = 0; }
    v->seq = p + 1 + r.l_hdr + 1; v->l_seq = (int)r.l_seq;
    v->qual = p + r.qual_rel; v->l_qual = (int)r.l_seq;
  }
  // records whose '@' lies in [pos, hi), blank lines between records skipped (kseq skips to the next '@' or '>')
  void scan_range(Chunk* c, size_t pos) const {
    const unsigned char* d = map;
    const size_t size = map_size;
    c->start = pos; c->stopped = false; c->recs.clear();
    for (;;) {
      size_t q = pos;
      while (q < size && d[q] == '\n') ++q;
      if (q >= size) { c->end_pos = size; return; }
      if (q >= c->hi) { c->end_pos = q; return; }
      B2Rec r;
      size_t next;
      if (!parse_rec(d, size, q, &r, &next)) { c->end_pos = q; c->stopped = true; return; }
      c->recs.push_back(r);
      pos = next;
    }
  }
  // first position in [lo, hi) that looks like the start of a record (kNoStart: none)
  size_t find_start(size_t lo, size_t hi) const {
    const unsigned char* d = map;
    const size_t size = map_size;
    size_t p = lo;
    if (lo > 0 && d[lo - 1] != '\n') {
      const unsigned char* nl = (const unsigned char*)memchr(d + lo, '\n', size - lo);
      if (!nl) return kNoStart;
      p = (size_t)(nl - d) + 1;
    }
    for (int tries = 0; tries < 256; ++tries) {
      while (p < size && d[p] == '\n') ++p;
      if (p >= hi || p >= size) return kNoStart;
      B2Rec r;
      size_t next;
      if (d[p] == '@' && parse_rec(d, size, p, &r, &next)) {
        size_t q = next;
        while (q < size && d[q] == '\n') ++q;
        size_t n2;
        if (q >= size || parse_rec(d, size, q, &r, &n2)) return p;
      }
      const unsigned char* nl = (const unsigned char*)memchr(d + p, '\n', size - p);
      if (!nl) return kNoStart;
      p = (size_t)(nl - d) + 1;
    }
    return kNoStart;
  }
  void pre_start() {
    if (!map || pre_on) return;
    if (const char* s = getenv("B200BWA_SCAN_CHUNK")) { long v = atol(s); if (v >= 64) chunk_bytes = (size_t)v; }
    const size_t n_chunks = (map_size + chunk_bytes - 1) / chunk_bytes;
    // a quarter of the host cores per input file, between 2 and 8 (the scan runs at ~2 GB/s per thread; one job over
    // 8 GPUs consumes FASTQ at 10+ GB/s)
    int n_threads = (int)std::thread::hardware_concurrency() / 4;
    n_threads = n_threads < 2 ? 2 : n_threads > 8 ? 8 : n_threads;
    if (const char* s = getenv("B200BWA_SCAN_THREADS")) n_threads = atoi(s);
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_chunks) n_threads = (int)n_chunks;
    chunks.clear();
    chunks.resize(n_chunks);
    pre_window = (size_t)2 * n_threads + 2;
    pre_next_chunk = 0; pre_consumed = 0; pre_stop = false; pre_on = true;
    cur_chunk = 0; cur_i = 0; cur_expected = 0; cur_ready = false;
    for (int t = 0; t < n_threads; ++t)
      pre_threads.emplace_back([this, n_chunks]() {
        for (;;) {
          size_t k;
          {
            std::unique_lock<std::mutex> lk(pre_mu);
            pre_cv.wait(lk, [&] { return pre_stop || pre_next_chunk >= n_chunks || pre_next_chunk < pre_consumed + pre_window; });
            if (pre_stop || pre_next_chunk >= n_chunks) return;
            k = pre_next_chunk++;
          }
          std::unique_ptr<Chunk> c(new Chunk);
          c->lo = k * chunk_bytes;
          c->hi = std::min(map_size, c->lo + chunk_bytes);
          c->recs.reserve((c->hi - c->lo) / 256 + 16);
          const size_t st = k == 0 ? 0 : find_start(c->lo, c->hi);
          if (st == kNoStart) { c->start = kNoStart; c->end_pos = kNoStart; }
          else scan_range(c.get(), st);
          c->done = true;
          std::lock_guard<std::mutex> lk(pre_mu);
          chunks[k] = std::move(c);
          pre_cv.notify_all();
        }
      });
  }
  // Makes chunk cur_chunk current and validated.  False when the scan-ahead has ended (EOF, or an irregular record at
  // `begin`, which then holds the position to continue from).
  bool pre_load() {
    for (;;) {
      if (cur_ready) {
        Chunk* c = chunks[cur_chunk].get();
        if (cur_i < c->recs.size()) return true;
        // chunk used up
        const size_t endp = c->end_pos;
        const bool stopped = c->stopped;
        {
          std::lock_guard<std::mutex> lk(pre_mu);
          chunks[cur_chunk].reset();
          pre_consumed = cur_chunk + 1;
          pre_cv.notify_all();
        }
        cur_expected = endp;
        cur_ready = false;
        ++cur_chunk;
        if (stopped || cur_chunk >= chunks.size() || endp >= map_size) { pre_finish(endp); return false; }
      }
      {
        std::unique_lock<std::mutex> lk(pre_mu);
        pre_cv.wait(lk, [&] { return chunks[cur_chunk] != nullptr; });
      }
      Chunk* c = chunks[cur_chunk].get();
      if (cur_expected >= c->hi) {           // the previous chunk's last record covers this chunk entirely
        c->recs.clear(); c->end_pos = cur_expected; c->stopped = false;
      } else if (c->start != cur_expected) {  // speculation missed: scan it here from the right position
        scan_range(c, cur_expected);
      }
      cur_i = 0;
      cur_ready = true;
    }
  }
  void pre_finish(size_t pos) {
    pre_shutdown();
    begin = pos;
  }
  bool pre_span(const B2Rec** p, size_t* n) {
    if (!pre_on || !pre_load()) return false;
    Chunk* c = chunks[cur_chunk].get();
    *p = c->recs.data() + cur_i; *n = c->recs.size() - cur_i;
    return true;
  }
  // next scanned-ahead view; false once the scan-ahead has ended and its records are used up
  bool pre_next(B2RecView* v) {
    const B2Rec* p;
    size_t n;
    if (!pre_span(&p, &n)) return false;
    rec_view(map, *p, v);
    ++cur_i;
    return true;
  }
  void pre_shutdown() {
    if (!pre_on) return;
    { std::lock_guard<std::mutex> lk(pre_mu); pre_stop = true; }
    pre_cv.notify_all();
    for (auto& t : pre_threads) t.join();
    pre_threads.clear();
    pre_on = false; chunks.clear(); cur_ready = false;
  }
  bool refill() {
    if (map || is_eof) { is_eof = true; return false; }
    begin = 0;
    int n = gzread(fp, store.data(), (unsigned)store.size());
    if (n < 0 && io_err.empty()) {
      int errnum = 0;
      const char* msg = gzerror(fp, &errnum);
      io_err = std::string("[gzread] ") + (errnum == Z_ERRNO ? strerror(errno) : (msg ? msg : "read error"));
    }
    end = n > 0 ? (size_t)n : 0;
    data = store.data();
    if (end == 0) { is_eof = true; return false; }
    return true;
  }
  int getc_() {
    if (begin >= end && !refill()) return -1;
    return data[begin++];
  }
  // delimiter: 0 = any whitespace, 2 = newline.  Returns length or -1 at EOF without data.
  int getuntil(int delim, std::string& str, int* dret, bool append) {
    bool gotany = false;
    if (dret) *dret = 0;
    if (!append) str.clear();
    for (;;) {
      if (begin >= end && !refill()) break;
      size_t i;
      if (delim == 2) {
        const unsigned char* q = (const unsigned char*)memchr(data + begin, '\n', end - begin);
        i = q ? (size_t)(q - data) : end;
      } else {
        for (i = begin; i < end; ++i) if (isspace(data[i])) break;
      }
      gotany = true;
      str.append((const char*)data + begin, i - begin);
      begin = i + 1;
      if (i < end) { if (dret) *dret = data[i]; break; }
    }
    if (!gotany && is_eof && begin >= end) return -1;
    if (delim == 2 && str.size() > 1 && str[str.size() - 1] == '\r') str.resize(str.size() - 1);
    return (int)str.size();
  }
  int read() {
    int c;
    if (last_char == 0) {
      while ((c = getc_()) != -1 && c != '>' && c != '@') {}
      if (c == -1) return -1;
      last_char = c;
    }
    comment.clear(); seq.clear(); qual.clear();
    if (getuntil(0, name, &c, false) < 0) return -1;
    if (c != '\n') getuntil(2, comment, 0, false);
    while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
      if (c == '\n') continue;
      seq += (char)c;
      getuntil(2, seq, 0, true);
    }
    if (c == '>' || c == '@') last_char = c;
    if (c != '+') return (int)seq.size();
    while ((c = getc_()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while (getuntil(2, qual, 0, true) >= 0 && qual.size() < seq.size()) {}
    last_char = 0;
    if (seq.size() != qual.size()) return -2;
    return (int)seq.size();
  }
  // plain 4-line record at the current position of the mapping?  (see the comment above)
  bool fast(B2RecView* v) {
    if (!map || last_char != 0) return false;
    size_t q = begin;
    while (q < end && data[q] == '\n') ++q;   // blank lines between records (SparkBWA writes one after every record)
    B2Rec r;
    size_t next;
    if (q >= end || !parse_rec(data, end, q, &r, &next)) return false;
    rec_view(data, r, v);
    begin = next;
    return true;
  }
};

B2SeqReader::B2SeqReader() : impl_(new Impl) {}
B2SeqReader::~B2SeqReader() { close(); delete impl_; }
int B2SeqReader::open(const char* fn) {
  close();
  delete impl_;
  impl_ = new Impl;
  if (strcmp(fn, "-") && !getenv("B200BWA_NO_MMAP")) {  // regular, non-empty, not gzip: map it
    int fd = ::open(fn, O_RDONLY);
    if (fd < 0) return 1;
    struct stat st;
    unsigned char magic[2] = {0, 0};
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 2 && pread(fd, magic, 2, 0) == 2 &&
        !(magic[0] == 0x1f && magic[1] == 0x8b)) {
      void* m = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        impl_->map = (const unsigned char*)m; impl_->map_size = (size_t)st.st_size;
        impl_->data = impl_->map; impl_->begin = 0; impl_->end = impl_->map_size;
      }
    }
    ::close(fd);
    if (impl_->map) {
      if (!getenv("B200BWA_NO_SCAN_AHEAD")) impl_->pre_start();
      return 0;
    }
  }
  impl_->fp = strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(0, "r");
  if (!impl_->fp) return 1;
  gzbuffer(impl_->fp, 1 << 20);
  impl_->store.resize(1 << 20);
  return 0;
}
void B2SeqReader::close() {
  impl_->pre_shutdown();
  if (impl_->fp) {   // a .gz cut off in mid-stream shows up here (Z_BUF_ERROR): fatal in the reference (err_gzclose, bwa/utils.c:284-289)
    const int r = gzclose(impl_->fp);
    impl_->fp = 0;
    if (r != Z_OK && impl_->io_err.empty()) impl_->io_err = std::string("[gzclose] ") + (r == Z_ERRNO ? strerror(errno) : zError(r));
  }
  if (impl_->map) { munmap((void*)impl_->map, impl_->map_size); impl_->map = 0; impl_->data = 0; impl_->begin = impl_->end = 0; }
}
int B2SeqReader::next() {  // string-returning form: same record stream as next_view
  B2RecView v;
  bool stable = false;
  const int r = next_view(&v, &stable);
  if (r >= 0 && stable) {
    impl_->name.assign(v.name, v.l_name); impl_->comment.assign(v.comment ? v.comment : "", v.l_comment);
    impl_->seq.assign(v.seq, v.l_seq); impl_->qual.assign(v.qual, v.l_qual);
  }
  return r;
}
int B2SeqReader::next_view(B2RecView* v, bool* stable) {
  if (impl_->pre_on) {
    if (impl_->pre_next(v)) { *stable = true; return v->l_seq; }
    // the scan-ahead thread has ended (EOF or a record for the general parser): carry on from its position
  }
  if (impl_->fast(v)) { *stable = true; return v->l_seq; }
  *stable = false;
  int r = impl_->read();
  if (r < 0) return r;
  v->name = impl_->name.data(); v->l_name = (int)impl_->name.size();
  v->comment = impl_->comment.data(); v->l_comment = (int)impl_->comment.size();
  v->seq = impl_->seq.data(); v->l_seq = (int)impl_->seq.size();
  v->qual = impl_->qual.data(); v->l_qual = (int)impl_->qual.size();
  return r;
}
bool B2SeqReader::span(const B2Rec** p, size_t* n) { return impl_->pre_span(p, n); }
void B2SeqReader::advance(size_t k) { impl_->cur_i += k; }
const unsigned char* B2SeqReader::map_base() const { return impl_->map; }
void B2SeqReader::view_of(const unsigned char* base, const B2Rec& r, B2RecView* v) { Impl::rec_view(base, r, v); }
const std::string& B2SeqReader::name() const { return impl_->name; }
const std::string& B2SeqReader::io_error() const { return impl_->io_err; }
const std::string& B2SeqReader::comment() const { return impl_->comment; }
const std::string& B2SeqReader::seq() const { return impl_->seq; }
const std::string& B2SeqReader::qual() const { return impl_->qual; }

char* B2RawBatch::hold(const char* p, size_t n) {
  if (n == 0) return 0;
  if (arena.empty() || arena_used + n > arena_cap) {
    arena_cap = n > ((size_t)4 << 20) ? n : ((size_t)4 << 20);
    arena.emplace_back(new char[arena_cap]);
    arena_used = 0;
  }
  char* d = arena.back().get() + arena_used;
  memcpy(d, p, n);
  arena_used += n;
  return d;
}

static inline int take_record(B2SeqReader* ks, B2RawBatch* rb) {
  B2RecView v;
  bool stable = false;
  if (ks->next_view(&v, &stable) < 0) return -1;
  if (!stable) {  // the reader's strings are overwritten by its next record
    v.name = rb->hold(v.name, v.l_name); v.comment = rb->hold(v.comment, v.l_comment);
    v.seq = rb->hold(v.seq, v.l_seq); v.qual = rb->hold(v.qual, v.l_qual);
  }
  // strlen semantics of the reference's strdup'ed sequence: cut at an embedded NUL (never on the mapped path)
  if (!stable && v.l_seq > 0) { const void* z = memchr(v.seq, 0, v.l_seq); if (z) v.l_seq = (int)((const char*)z - v.seq); }
  rb->recs.push_back(v);
  return v.l_seq;
}

// compact descriptors -> views, in read order (two files: interleaved)
static void expand_compact(B2RawBatch* rb) {
  if (!rb->compact) return;
  const size_t n0 = rb->crec[0].size(), n1 = rb->crec[1].size();
  rb->recs.reserve(n0 + n1 + 64);
  B2RecView v;
  for (size_t i = 0; i < n0; ++i) {
    B2SeqReader::view_of(rb->cbase[0], rb->crec[0][i], &v);
    rb->recs.push_back(v);
    if (i < n1) { B2SeqReader::view_of(rb->cbase[1], rb->crec[1][i], &v); rb->recs.push_back(v); }
  }
  rb->compact = false; rb->crec[0].clear(); rb->crec[1].clear();
}

// one batch exactly as bseq_read cuts it (bwa/bwa.c:42-75); returns the number of reads.  While every reader has
// scanned-ahead plain FASTQ records to offer the batch is collected as descriptors (rb->compact); the first record that
// needs the general parser turns it into views.
int b2_read_raw_batch(int chunk_size, B2SeqReader* ks, B2SeqReader* ks2, B2RawBatch* rb) {
  rb->clear();
  int size = 0, max_len = 0;
  bool full = false;
  if (!getenv("B200BWA_NO_COMPACT")) {
    rb->compact = true; rb->n_files = ks2 ? 2 : 1;
    rb->cbase[0] = ks->map_base(); rb->cbase[1] = ks2 ? ks2->map_base() : 0;
    while (!full) {
      const B2Rec *p1 = 0, *p2 = 0;
      size_t n1 = 0, n2 = 0;
      if (!ks->span(&p1, &n1)) break;
      if (ks2 && !ks2->span(&p2, &n2)) break;
      const size_t m = ks2 ? std::min(n1, n2) : n1;
      size_t i = 0;
      if (ks2) {
        while (i < m) {
          const int l1 = (int)p1[i].l_seq, l2 = (int)p2[i].l_seq;
          size += l1 + l2; ++i;
          if (l1 > max_len) max_len = l1;
          if (l2 > max_len) max_len = l2;
          if (size >= chunk_size) { full = true; break; }
        }
        rb->crec[1].insert(rb->crec[1].end(), p2, p2 + i);
        ks2->advance(i);
      } else {
        const size_t have = rb->crec[0].size();
        while (i < m) {
          const int l1 = (int)p1[i].l_seq;
          size += l1; ++i;
          if (l1 > max_len) max_len = l1;
          if (size >= chunk_size && ((have + i) & 1) == 0) { full = true; break; }
        }
      }
      rb->crec[0].insert(rb->crec[0].end(), p1, p1 + i);
      ks->advance(i);
    }
    if (rb->crec[0].empty()) rb->compact = false;
    // the device addresses the raw bytes of a batch with 32-bit offsets
    if (rb->compact) {
      uint64_t span = 0;
      for (int f = 0; f < rb->n_files; ++f) {
        const std::vector<B2Rec>& c = rb->crec[f];
        span += c.back().off + c.back().qual_rel + c.back().l_seq + 1 - c.front().off;
      }
      if (span >= ((uint64_t)1 << 31)) expand_compact(rb);
    }
  }
  if (!full) {
    for (;;) {
      B2RawBatch one;  // (records of the general parser are copied into the batch's arena by take_record)
      const size_t before = rb->recs.size();
      if (rb->compact) {
        // peek: does another record exist at all?  only then leave the compact form
        B2RawBatch* dst = &one;
        int l = take_record(ks, dst);
        if (l < 0) break;
        expand_compact(rb);
        B2RecView v = one.recs[0];
        if (!one.arena.empty()) {  // re-home the copies
          v.name = rb->hold(v.name, v.l_name); v.comment = rb->hold(v.comment, v.l_comment);
          v.seq = rb->hold(v.seq, v.l_seq); v.qual = rb->hold(v.qual, v.l_qual);
        }
        rb->recs.push_back(v);
        if (ks2) {
          int l2 = take_record(ks2, rb);
          if (l2 < 0) {
            fprintf(stderr, "[W::bseq_read] the 2nd file has fewer sequences.\n");
            rb->recs.pop_back();
            break;
          }
          size += l2;
        }
        size += l;
      } else {
        int l = take_record(ks, rb);
        if (l < 0) break;
        if (ks2) {
          int l2 = take_record(ks2, rb);
          if (l2 < 0) {
            fprintf(stderr, "[W::bseq_read] the 2nd file has fewer sequences.\n");
            rb->recs.pop_back();
            break;
          }
          size += l2;
        }
        size += l;
      }
      (void)before;
      if (size >= chunk_size && (rb->recs.size() & 1) == 0) break;
    }
  }
  if (size == 0) {
    B2RecView v; bool st;
    if (ks2 && ks2->next_view(&v, &st) >= 0) fprintf(stderr, "[W::bseq_read] the 1st file has fewer sequences.\n");
  }
  rb->n_bases = size;
  rb->max_len = max_len;
  return (int)rb->n_reads();
}

// record views -> the packed arrays the engine uploads (name trimming of bwa/bwa.c:27-31, nst_nt4_table codes)
void b2_pack_batch(const B2RawBatch& rb, int copy_comment, B2HostBatch* b) {
  // (no clear(): every array is resized to its exact length below, and a vector that only shrinks or regrows within
  // its old size is not zero-filled again -- 20 MB of memset per batch otherwise)
  const size_t n = rb.recs.size();
  size_t tot_seq = 0, tot_name = 0, tot_com = 0;
  for (const B2RecView& v : rb.recs) { tot_seq += v.l_seq; tot_name += v.l_name; tot_com += v.l_comment; }
  b->seq.resize(tot_seq); b->qual.resize(tot_seq); b->names.resize(tot_name);
  b->comments.resize(copy_comment ? tot_com : 0);
  b->has_qual.resize(n); b->seq_off.resize(n); b->name_off.resize(n); b->com_off.resize(n);
  b->l_seq.resize(n); b->l_name.resize(n); b->l_com.resize(n);
  size_t so = 0, no = 0, co = 0;
  int max_len = 0;
  for (size_t i = 0; i < n; ++i) {
    const B2RecView& v = rb.recs[i];
    size_t l = v.l_name ? strnlen(v.name, v.l_name) : 0;
    if (l > 2 && v.name[l - 2] == '/' && isdigit((unsigned char)v.name[l - 1])) l -= 2;  // trim_readno
    const size_t ls = v.l_seq;
    b->seq_off[i] = (int64_t)so;
    b->l_seq[i] = (int32_t)ls;
    uint8_t* ds = b->seq.data() + so;
    for (size_t k = 0; k < ls; ++k) ds[k] = nt4_table[(unsigned char)v.seq[k]];
    if (v.l_qual > 0) { memcpy(b->qual.data() + so, v.qual, ls); b->has_qual[i] = 1; }
    else { memset(b->qual.data() + so, 0, ls); b->has_qual[i] = 0; }
    so += ls;
    b->name_off[i] = (int64_t)no;
    b->l_name[i] = (int32_t)l;
    if (l) memcpy(b->names.data() + no, v.name, l);
    no += l;
    if (copy_comment && v.l_comment > 0) {
      b->com_off[i] = (int64_t)co;
      b->l_com[i] = v.l_comment;
      memcpy(b->comments.data() + co, v.comment, v.l_comment);
      co += v.l_comment;
    } else { b->com_off[i] = 0; b->l_com[i] = -1; }
    if ((int)ls > max_len) max_len = (int)ls;
  }
  b->seq.resize(so); b->qual.resize(so); b->names.resize(no); b->comments.resize(co);
  b->n_reads = (int)n; b->n_bases = (int64_t)so; b->max_len = max_len;
  b->n_processed = rb.n_processed;
}

int b2_read_batch(int chunk_size, B2SeqReader* ks, B2SeqReader* ks2, int copy_comment, B2HostBatch* b) {
  B2RawBatch rb;
  int n = b2_read_raw_batch(chunk_size, ks, ks2, &rb);
  b2_pack_batch(rb, copy_comment, b);
  return n;
}

// ---------------------------------------------------------------------------------------------
// Index cache: like `bwa shm` keeps the index in host shared memory (bwa/bwashm.c:87-118), the
// device images stay resident across calls in one executor process.  An entry is keyed by the index
// prefix plus the size and mtime of its .bwt file (a rebuilt index under the same name is reloaded)
// and holds one engine per device it has been used on; the first engine is filled from the files,
// every further one by a device-to-device copy.  A call holds the entry's lock while it runs: calls on
// different indices run concurrently, calls on the same index take turns (they share the engines'
// stream/buffer sets).
struct B2CachedIndex {
  std::string prefix;
  long long bwt_size = 0, bwt_mtime_ns = 0;
  bool ignore_alt = false;
  std::unique_ptr<B2IndexHost> host;
  std::vector<std::unique_ptr<B2Engine>> engines;  // one per device, in order of first use
  // Engines are leased to calls: a call takes every engine of its device list that is free (at least one, at most
  // B200BWA_DEVICES_PER_CALL), so several tasks of one executor run side by side on different GPUs while a lone task
  // spreads over all of them.
  std::mutex lease_mu;
  std::condition_variable lease_cv;
  std::vector<B2Engine*> leased;
  // the ord-th engine on device d (a device listed twice in B200BWA_DEVICES gets two engines: that is how the
  // multi-engine path is exercised on a box with a single GPU)
  B2Engine* on_device(int d, int ord = 0) {
    for (auto& e : engines) if (e->device() == d && ord-- == 0) return e.get();
    return 0;
  }
};
static std::mutex g_cache_mu;
static std::vector<std::shared_ptr<B2CachedIndex>> g_cache;

static bool stat_bwt(const char* hint, long long* size, long long* mtime_ns) {
  std::string prefix = infer_prefix(hint);
  struct stat st;
  if (prefix.empty() || stat((prefix + ".bwt").c_str(), &st) != 0) return false;
  *size = (long long)st.st_size;
  *mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
  return true;
}

// engines for `devices` (all on one index), created on first use; null on failure
static std::shared_ptr<B2CachedIndex> b2_get_engines(const char* prefix, const std::vector<int>& devices, int ignore_alt,
                                                     int verbose, std::vector<B2Engine*>* out) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  long long sz = 0, mt = 0;
  if (!stat_bwt(prefix, &sz, &mt)) {
    if (verbose >= 1) fprintf(stderr, "[E::bwa_idx_load_from_disk] fail to locate the index files\n");
    set_err("fail to locate the index files for '%s'", prefix);
    return nullptr;
  }
  std::shared_ptr<B2CachedIndex> c;
  for (size_t i = 0; i < g_cache.size(); ++i) {
    B2CachedIndex& e = *g_cache[i];
    if (e.prefix != prefix || e.ignore_alt != (ignore_alt != 0)) continue;
    if (e.bwt_size != sz || e.bwt_mtime_ns != mt) {  // the files changed under this name: forget the stale image
      g_cache.erase(g_cache.begin() + i);
      break;
    }
    c = g_cache[i];
    break;
  }
  B2Opt opt0;
  b2_opt_init(&opt0);
  bool fresh = false;
  if (!c) {
    c.reset(new B2CachedIndex);
    c->prefix = prefix; c->bwt_size = sz; c->bwt_mtime_ns = mt; c->ignore_alt = ignore_alt != 0;
    c->host.reset(new B2IndexHost);
    fresh = true;
  }
  for (size_t di = 0; di < devices.size(); ++di) {
    const int d = devices[di];
    int ord = 0;
    for (size_t k = 0; k < di; ++k) if (devices[k] == d) ++ord;
    if (c->on_device(d, ord)) continue;
    std::unique_ptr<B2Engine> e(new B2Engine);
    B2EngineConfig cfg;
    cfg.verbose = verbose;
    b2_engine_config_from_env(&cfg);
    cfg.device = d;
    if (c->engines.empty()) {
      if (b2_load_index(prefix, c->host.get(), verbose, 0)) return nullptr;
      if (ignore_alt) for (auto& a : c->host->anns) a.is_alt = 0;
      if (e->init(*c->host, opt0, "", cfg)) { set_err("%s", e->last_error()); return nullptr; }
      // the host copies of the big arrays are no longer needed once they are in HBM
      std::vector<uint32_t>().swap(c->host->bwt);
      std::vector<uint64_t>().swap(c->host->sa);
      std::vector<uint8_t>().swap(c->host->pac);
    } else {
      // fan out: copy from the engine that was created log2 steps ago (0 -> 1, then 0 -> 2 and 1 -> 3, ...), so the
      // sources of successive copies are different GPUs (the copies themselves are queued per destination device)
      size_t k = c->engines.size(), p2 = 1;
      while (p2 * 2 <= k) p2 *= 2;
      if (e->init_from_peer(*c->host, *c->engines[k - p2], opt0, "", cfg)) { set_err("%s", e->last_error()); return nullptr; }
    }
    c->engines.push_back(std::move(e));
  }
  if (fresh) g_cache.push_back(c);
  if (out) {
    out->clear();
    for (size_t di = 0; di < devices.size(); ++di) {
      int ord = 0;
      for (size_t k = 0; k < di; ++k) if (devices[k] == devices[di]) ++ord;
      out->push_back(c->on_device(devices[di], ord));
    }
  }
  return c;
}

B2Engine* b2_get_engine(const char* prefix, int device, int ignore_alt, const B2Opt& opt, const char* rg_id, int verbose,
                        B2IndexHost** host_out) {
  std::vector<B2Engine*> engs;
  std::shared_ptr<B2CachedIndex> c = b2_get_engines(prefix, std::vector<int>(1, device), ignore_alt, verbose, &engs);
  if (!c) return 0;
  B2Engine* e = engs[0];
  e->set_options(opt, rg_id, verbose);
  if (host_out) *host_out = c->host.get();
  return e;
}

// drops every cached device image that no call is using (the executor is about to load a different index, or wants
// its HBM back); returns the number of entries released
int b2_index_cache_evict() {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  int n = 0;
  for (size_t i = 0; i < g_cache.size();) {
    bool idle = g_cache[i].use_count() == 1;
    if (idle) { std::lock_guard<std::mutex> l2(g_cache[i]->lease_mu); idle = g_cache[i]->leased.empty(); }
    if (idle) {
      g_cache.erase(g_cache.begin() + i);
      ++n;
    } else ++i;
  }
  return n;
}

void b2_engine_config_from_env(B2EngineConfig* cfg) {
  const char* s;
  if ((s = getenv("B200BWA_LANE_GRID"))) cfg->lane_grid = atoi(s);
  if ((s = getenv("B200BWA_LANE_BLOCK"))) cfg->lane_block = atoi(s);
  if ((s = getenv("B200BWA_SEED_GRID"))) cfg->seed_grid = atoi(s);
  if ((s = getenv("B200BWA_SEED_BLOCK"))) cfg->seed_block = atoi(s);
  if ((s = getenv("B200BWA_SEED_ENTRIES"))) cfg->seed_smem_entries = atoi(s);
  if ((s = getenv("B200BWA_SEED_BPS"))) cfg->seed_blocks_per_sm = atoi(s);
  if ((s = getenv("B200BWA_SEED3_BPS"))) cfg->seed3_blocks_per_sm = atoi(s);
  if ((s = getenv("B200BWA_FINAL_HEAVY"))) cfg->final_heavy_regs = atoi(s);
  if ((s = getenv("B200BWA_FINAL_FAST"))) cfg->final_fast_bytes = atoi(s);
  if ((s = getenv("B200BWA_SCRATCH"))) cfg->scratch_per_lane = (size_t)atol(s);
  if ((s = getenv("B200BWA_BIG_POOL"))) cfg->big_pool_bytes = (size_t)atoll(s);
  if ((s = getenv("B200BWA_BIG_SLOT"))) cfg->big_slot_bytes = (size_t)atoll(s);
  if ((s = getenv("B200BWA_MAX_SCRATCH"))) cfg->max_scratch_bytes = (size_t)atoll(s);
  if ((s = getenv("B200BWA_DEVICE"))) cfg->device = atoi(s);
  if ((s = getenv("B200BWA_WORKERS"))) cfg->n_workers = atoi(s);
  if ((s = getenv("B200BWA_CAP_INTV"))) cfg->cap_intv = atoi(s);
  if ((s = getenv("B200BWA_SLOT"))) cfg->slot_per_read = atoi(s);
  if ((s = getenv("B200BWA_NO_GALN"))) cfg->no_galn = atoi(s);
  if ((s = getenv("B200BWA_NO_SA_PLAN"))) cfg->no_sa_plan = atoi(s);
  if ((s = getenv("B200BWA_NO_XEXT"))) cfg->no_xext = atoi(s);
  if ((s = getenv("B200BWA_NO_EXT_CHAIN"))) cfg->no_ext_chain = atoi(s);
  if ((s = getenv("B200BWA_NO_RESC_PLAN"))) cfg->no_resc_plan = atoi(s);
}

// ---------------------------------------------------------------------------------------------
// main_mem
namespace {
struct Job {
  int64_t index = 0;
  int n_reads = 0;
  std::unique_ptr<B2RawBatch> raw;
  std::string sam;           // ordered-writer mode only
  size_t sam_size = 0;
  int64_t offset = -1;       // positioned-write mode: file offset of this batch's SAM block
  bool size_known = false;
  const char* wdata = 0;     // positioned-write mode: the text (in a pinned buffer of the worker, or in `sam`)
  int wslot = -1;            // which of the worker's pinned buffers (index into `busy`), -1: none
  double t_gpu = 0;
  int rc = 0;
  bool done = false;
};
}  // namespace

// bseq_classify (bwa/bwa.c:77-93): adjacent reads with equal names form a pair, everything else is single-end
static void classify_smart(const B2HostBatch& b, std::vector<int>& se, std::vector<int>& pe) {
  const int n = b.n_reads;
  auto same = [&](int x, int y) {
    return b.l_name[x] == b.l_name[y] && memcmp(b.names.data() + b.name_off[x], b.names.data() + b.name_off[y], b.l_name[x]) == 0;
  };
  int i, has_last;
  for (i = 1, has_last = 1; i < n; ++i) {
    if (has_last) {
      if (same(i, i - 1)) { pe.push_back(i - 1); pe.push_back(i); has_last = 0; }
      else se.push_back(i - 1);
    } else has_last = 1;
  }
  if (has_last && n > 0) se.push_back(i - 1);
}

static void subset_batch(const B2HostBatch& src, const std::vector<int>& idx, B2HostBatch* dst) {
  dst->clear();
  for (int r : idx) {
    const size_t ls = src.l_seq[r];
    dst->seq_off.push_back((int64_t)dst->seq.size());
    dst->l_seq.push_back(src.l_seq[r]);
    dst->seq.insert(dst->seq.end(), src.seq.begin() + src.seq_off[r], src.seq.begin() + src.seq_off[r] + ls);
    dst->qual.insert(dst->qual.end(), src.qual.begin() + src.seq_off[r], src.qual.begin() + src.seq_off[r] + ls);
    dst->has_qual.push_back(src.has_qual[r]);
    dst->name_off.push_back((int64_t)dst->names.size());
    dst->l_name.push_back(src.l_name[r]);
    dst->names.insert(dst->names.end(), src.names.begin() + src.name_off[r], src.names.begin() + src.name_off[r] + src.l_name[r]);
    if (src.l_com[r] >= 0) {
      dst->com_off.push_back((int64_t)dst->comments.size());
      dst->l_com.push_back(src.l_com[r]);
      dst->comments.insert(dst->comments.end(), src.comments.begin() + src.com_off[r], src.comments.begin() + src.com_off[r] + src.l_com[r]);
    } else { dst->com_off.push_back(0); dst->l_com.push_back(-1); }
    ++dst->n_reads;
    dst->n_bases += (int64_t)ls;
    if ((int)ls > dst->max_len) dst->max_len = (int)ls;
  }
}

// One smart-pairing batch (bwa/fastmap.c:64-83): the single-end reads first (ids from n_processed), then the pairs
// (ids from n_processed + #single), SAM records returned in the input order.
static int process_smart(B2Engine* eng, const B2HostBatch& hb, const B2Pes* pes0, int wi, int verbose, std::string& out) {
  std::vector<int> se, pe;
  classify_smart(hb, se, pe);
  if (verbose >= 3) fprintf(stderr, "[M::process] %d single-end sequences; %d paired-end sequences\n", (int)se.size(), (int)pe.size());
  B2HostBatch sub;
  std::string sam[2];
  std::vector<int32_t> len[2];
  if (!se.empty()) {
    subset_batch(hb, se, &sub);
    sub.n_processed = hb.n_processed;
    if (eng->process_items(sub, 0, wi, 0, sam[0], len[0])) return 1;
  }
  if (!pe.empty()) {
    subset_batch(hb, pe, &sub);
    sub.n_processed = hb.n_processed + (int64_t)se.size();
    if (eng->process_items(sub, pes0, wi, 1, sam[1], len[1])) return 1;
  }
  // merge: walk the reads in input order; a pair's block (both mates) sits at its first read
  size_t is = 0, ip = 0, os = 0, op = 0;
  for (int i = 0; i < hb.n_reads;) {
    if (ip < pe.size() && pe[ip] == i) {
      out.append(sam[1], op, (size_t)len[1][ip >> 1]); op += (size_t)len[1][ip >> 1];
      ip += 2; i += 2;
    } else {
      out.append(sam[0], os, (size_t)len[0][is]); os += (size_t)len[0][is];
      ++is; ++i;
    }
  }
  return 0;
}

int b2_mem_main(int argc, char** argv, const char* out_path, const char* pg_cmdline) {
  B2MemArgs m;
  double t_real = b2_realtime();
  if (b2_parse_mem_args(argc, argv, &m)) return 1;
  if (m.positional.size() < 2 || m.positional.size() > 3) {
    fprintf(stderr, "\nUsage: bwa mem [options] <idxbase> <in1.fq> [in2.fq]\n\n");
    fprintf(stderr, "b200bwa accepts the option set of bwa-0.7.15 `mem` (see bwa.1).\n\n");
    return 1;
  }
  B2Opt opt = m.opt;
  B2SeqReader ks, ks2;
  bool have2 = false;
  const int chunk = m.fixed_chunk_size > 0 ? m.fixed_chunk_size : opt.chunk_size * opt.n_threads;
  // Devices of this call (the -K batches of ONE job are dealt over the GPUs of the box, batch b to
  // device b mod G, each carrying its absolute n_processed; one ordered output).  B200BWA_DEVICES = "all" or a
  // comma-separated list; B200BWA_DEVICE = one device; default: every visible device, but no more than the input
  // can keep busy (estimated from the file sizes: about 45 % of a FASTQ file's bytes are bases).
  std::vector<int> devices;
  {
    const char* ds = getenv("B200BWA_DEVICES");
    const int n_vis = b2_device_count();
    if (ds && strcmp(ds, "all") != 0) {
      for (const char* p = ds; *p;) {
        char* e;
        long v = strtol(p, &e, 10);
        if (e == p) break;
        devices.push_back((int)v);
        p = *e == ',' ? e + 1 : e;
      }
    } else if (!ds && getenv("B200BWA_DEVICE")) devices.push_back(atoi(getenv("B200BWA_DEVICE")));
    else {
      double bases = 0;
      bool known = true;
      for (size_t i = 1; i < m.positional.size(); ++i) {
        struct stat st;
        if (stat(m.positional[i].c_str(), &st) == 0 && S_ISREG(st.st_mode)) bases += 0.45 * (double)st.st_size;
        else known = false;
      }
      int want = n_vis < 1 ? 1 : n_vis;
      if (known && !ds) { double nb = ceil(bases / (double)(chunk > 0 ? chunk : 1)); if (nb < want) want = nb < 1 ? 1 : (int)nb; }
      for (int d = 0; d < want; ++d) devices.push_back(d);
    }
    if (devices.empty()) devices.push_back(0);
  }
  // index first (same order of failure modes as the reference: index, then reads)
  std::vector<B2Engine*> engs;
  std::shared_ptr<B2CachedIndex> cached;
  auto get_engines = [&]() -> bool {
    cached = b2_get_engines(m.positional[0].c_str(), devices, m.ignore_alt, m.verbose, &engs);
    // (a missing index has had its reference-style message already; a truncated or inconsistent one -- where the reference
    // aborts or crashes -- is named here)
    if (!cached && m.verbose >= 1 && b2_tls_error.find("fail to locate") == std::string::npos)
      fprintf(stderr, "[E::b200bwa] %s\n", b2_tls_error.c_str());
    return cached != nullptr;
  };
  if (ks.open(m.positional[1].c_str())) {
    // the reference loads the index before opening the reads; keep its message and code
    if (!get_engines()) return 1;
    if (m.verbose >= 1) fprintf(stderr, "[E::main_mem] fail to open file `%s'.\n", m.positional[1].c_str());
    return set_err("fail to open file `%s'", m.positional[1].c_str());
  }
  if (m.positional.size() == 3) {
    if (opt.flag & B2_F_PE) {
      if (m.verbose >= 2) fprintf(stderr, "[W::main_mem] when '-p' is in use, the second query file is ignored.\n");
    } else {
      if (ks2.open(m.positional[2].c_str())) {
        if (m.verbose >= 1) fprintf(stderr, "[E::main_mem] fail to open file `%s'.\n", m.positional[2].c_str());
        return set_err("fail to open file `%s'", m.positional[2].c_str());
      }
      have2 = true;
      opt.flag |= B2_F_PE;
    }
  }
  const bool smart = (opt.flag & B2_F_SMARTPE) != 0;  // interleaved input: pairs are adjacent reads with equal names
  const bool times0 = getenv("B200BWA_TIMES") != 0;
  if (times0) fprintf(stderr, "[T::b200bwa] inputs open at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  if (!get_engines()) return 1;
  if (times0) fprintf(stderr, "[T::b200bwa] engines ready at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  // Lease engines: every free one of the list (at least one: wait for it), at most B200BWA_DEVICES_PER_CALL.  Calls on
  // other indices proceed independently; calls on this index that find no free engine take turns.
  struct Lease {
    B2CachedIndex* c; std::vector<B2Engine*> mine;
    ~Lease() { std::lock_guard<std::mutex> lk(c->lease_mu); for (B2Engine* e : mine) c->leased.erase(std::find(c->leased.begin(), c->leased.end(), e)); c->lease_cv.notify_all(); }
  } lease{cached.get(), {}};
  {
    size_t per_call = engs.size();
    if (const char* pc = getenv("B200BWA_DEVICES_PER_CALL")) { long v = atol(pc); if (v >= 1 && (size_t)v < per_call) per_call = (size_t)v; }
    std::unique_lock<std::mutex> lk(cached->lease_mu);
    for (;;) {
      for (B2Engine* e : engs)
        if (lease.mine.size() < per_call && std::find(cached->leased.begin(), cached->leased.end(), e) == cached->leased.end() &&
            std::find(lease.mine.begin(), lease.mine.end(), e) == lease.mine.end())
          lease.mine.push_back(e);
      if (!lease.mine.empty()) break;
      cached->lease_cv.wait(lk);
    }
    for (B2Engine* e : lease.mine) cached->leased.push_back(e);
    engs = lease.mine;
  }
  const int G = (int)engs.size();
  B2IndexHost* host = cached->host.get();
  for (B2Engine* e : engs) e->set_options(opt, m.rg_id.c_str(), m.verbose);

  // Output.  A regular file (what SparkBWA passes with -f) is written with positioned writes issued by the GPU
  // worker threads straight from their pinned D2H buffers: a batch's offset is the header plus the sizes of all
  // earlier batches, known as soon as those have left the GPU.  Anything else (stdout, pipes) goes through the
  // ordered writer below.
  FILE* out = stdout;
  int out_fd = -1;
  if (out_path) {
    out_fd = ::open(out_path, O_RDWR | O_CREAT | O_TRUNC, 0666);
    if (out_fd < 0) { fprintf(stderr, "[E::main_mem] fail to open output `%s'.\n", out_path); return set_err("fail to open output `%s'", out_path); }
    struct stat st;
    if (fstat(out_fd, &st) != 0 || !S_ISREG(st.st_mode) || getenv("B200BWA_ORDERED_WRITER")) {
      out = fdopen(out_fd, "wb");
      if (!out) { ::close(out_fd); return set_err("fail to open output `%s'", out_path); }
      setvbuf(out, 0, _IOFBF, 1 << 22);
      out_fd = -1;
    } else out = 0;
  }
  const bool positioned = out_fd >= 0;
  auto pwrite_all = [&](const char* p, size_t n, int64_t off) -> bool {
    while (n > 0) {
      ssize_t k = ::pwrite(out_fd, p, n, (off_t)off);
      if (k < 0) { if (errno == EINTR) continue; return false; }
      p += k; n -= (size_t)k; off += k;
    }
    return true;
  };
  // Multi-GPU sharding: B200BWA_SHARD=<rank>/<world> makes this call align only the -K batches
  // b with b % world == rank (each still tagged with its absolute n_processed).  Rank 0 writes the header; every
  // rank writes "<out>.batches" (batch index, reads, SAM bytes per line) so the pieces can be merged in order.
  int shard_rank = 0, shard_world = 1;
  if (const char* sh = getenv("B200BWA_SHARD")) {
    if (sscanf(sh, "%d/%d", &shard_rank, &shard_world) != 2 || shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world) {
      if (out_path) { if (out) fclose(out); else ::close(out_fd); }
      return set_err("bad B200BWA_SHARD '%s' (want rank/world)", sh);
    }
  }
  FILE* fidx = 0;
  if (shard_world > 1 && out_path) fidx = fopen((std::string(out_path) + ".batches").c_str(), "w");
  std::string hdr = b2_sam_header(*host, m.hdr_line, pg_cmdline ? pg_cmdline : "");
  int64_t next_off = 0;  // positioned mode: where the next unassigned batch goes
  int64_t hdr_bytes = 0; // length of the header: what a failed run leaves behind
  bool write_failed = false;
  if (shard_rank == 0) {
    if (positioned) { if (!pwrite_all(hdr.data(), hdr.size(), 0)) write_failed = true; next_off = hdr_bytes = (int64_t)hdr.size(); }
    else fwrite(hdr.data(), 1, hdr.size(), out);
  }

  const B2Pes* pes0 = m.has_pes0 ? m.pes0 : 0;
  const int n_workers = engs[0]->n_workers();

  // Pipeline (kt_pipeline's three ordered steps): reader thread (record boundaries only) -> GPU workers (pack,
  // upload, kernels, download, positioned write) -> this thread (ordered bookkeeping / ordered writer).
  std::mutex mu;
  std::condition_variable cv;
  std::vector<std::deque<std::shared_ptr<Job>>> todo(G);  // one queue per device
  std::deque<std::shared_ptr<Job>> inorder;
  std::deque<std::shared_ptr<Job>> unassigned;  // positioned mode: jobs in order whose offset is not yet known
  bool read_done = false, failed = write_failed;
  std::string errmsg;
  const size_t max_inflight = (size_t)G * ((size_t)n_workers + 2);

  // One mutex, one condition variable per role (reader, each device's workers, writers, allocator, the ordering loop
  // below): with 48 workers and 24 writers on 8 GPUs a single notify_all would wake ~80 threads for every event.
  std::condition_variable cv_r, cv_w, cv_a, cv_m;
  std::vector<std::condition_variable> cv_dev(G);
  auto wake_all = [&]() { cv.notify_all(); cv_r.notify_all(); cv_w.notify_all(); cv_a.notify_all(); cv_m.notify_all(); for (auto& x : cv_dev) x.notify_all(); };
  std::thread reader([&]() {
    // B200BWA_FIRST_READ: absolute index of this input's first read within a larger run (a partition of a job that
    // was split upstream); it enters the hash tie-breaks exactly like n_processed of bwa/fastmap.c:85
    int64_t n_processed = getenv("B200BWA_FIRST_READ") ? atoll(getenv("B200BWA_FIRST_READ")) : 0;
    int64_t batch_index = 0, dealt = 0;
    // The reference's pipeline has two threads that take turns at the reading step (one with -1), and each leaves when a
    // read attempt of its own returns no sequences (bwa/kthread.c:92-102, bwa/fastmap.c:44-47): the input is at its end
    // after the SECOND empty attempt.  Normally both happen at EOF; but a malformed record (kseq_read returns -2: a
    // quality string longer or shorter than its sequence) ends the batch being read, and when it is the first record of
    // an attempt that attempt is empty -- the reference then carries on behind it once (bseq_read resumes at the next
    // '@' or '>'), and stops for good at the second such record or at EOF.
    int empty_attempts_left = m.no_mt_io ? 1 : 2;
    for (;;) {
      std::shared_ptr<Job> j(new Job);
      j->raw.reset(new B2RawBatch);
      int n = b2_read_raw_batch(chunk, &ks, have2 ? &ks2 : 0, j->raw.get());
      if (n == 0) { if (--empty_attempts_left == 0) break; continue; }
      j->raw->n_processed = n_processed;
      j->index = batch_index;
      j->n_reads = n;
      n_processed += n;
      if (batch_index++ % shard_world != shard_rank) continue;
      if (m.verbose >= 3) fprintf(stderr, "[M::process] read %d sequences (%ld bp)...\n", n, (long)j->raw->n_bases);
      if (times0) fprintf(stderr, "[T::b200bwa] batch %ld cut at %.1f ms\n", (long)j->index, (b2_realtime() - t_real) * 1e3);
      std::unique_lock<std::mutex> lk(mu);
      cv_r.wait(lk, [&] { return inorder.size() < max_inflight || failed; });
      if (failed) break;
      const int to_dev = (int)(dealt++ % G);
      todo[to_dev].push_back(j);
      inorder.push_back(j);
      if (positioned) unassigned.push_back(j);
      cv_dev[to_dev].notify_all();
    }
    std::lock_guard<std::mutex> lk(mu);
    {  // a read error is fatal in the reference (err_gzread exits): the run fails, the output keeps its header only
      const std::string& e1 = ks.io_error();
      const std::string& e2 = have2 ? ks2.io_error() : e1;
      if (!e1.empty() || !e2.empty()) { failed = true; if (errmsg.empty()) errmsg = !e1.empty() ? e1 : e2; }
    }
    read_done = true;
    wake_all();
  });

  // Positioned mode: the file writes are done by a few writer threads, so that a GPU worker starts its next batch
  // as soon as the SAM text of the current one sits in its pinned buffer (each worker has two, used in turn).
  // write()/pwrite() on ONE file are serialised by the kernel's inode lock (about 3 GB/s on tmpfs however many
  // threads call it: 8 M reads/s of SAM, less than two GPUs produce), so the writers copy through a shared mapping
  // of the output file instead, into pages an allocator thread has reserved ahead of them with fallocate() -- a full
  // disk then shows up as ENOSPC from fallocate, never as SIGBUS inside the executor.  File systems without
  // fallocate keep the positioned pwrite().
  char* out_map = 0;
  const size_t out_map_len = (size_t)1 << 38;
  int64_t alloc_end = 0;        // the allocator thread's frontier
  int64_t file_size = 0;        // size the file has been extended to (offsets handed out never exceed it)
  bool alloc_stop = false, populate_ok = false;
  std::thread allocator;
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
  if (positioned && !getenv("B200BWA_NO_MMAP_OUT") && fallocate(out_fd, 0, 0, (off_t)(next_off > 0 ? next_off : 1)) == 0) {
    void* mp = mmap(0, out_map_len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, out_fd, 0);
    if (mp != MAP_FAILED) {
      out_map = (char*)mp;
      alloc_end = file_size = next_off > 0 ? next_off : 1;
      // Writers reserve their own range with MADV_POPULATE_WRITE (Linux >= 5.14: the page faults of a write, batched,
      // with an error return where a store would raise SIGBUS) -- allocation then runs in parallel in every writer --
      // while the allocator thread works ahead of the offsets handed out so far.  Without that madvise the writers
      // wait for the allocator.
      populate_ok = !getenv("B200BWA_NO_POPULATE") && madvise(out_map, 4096, MADV_POPULATE_WRITE) == 0;
      const int64_t look = ((int64_t)16 << 20) * G + ((int64_t)64 << 20), step = (int64_t)16 << 20;
      allocator = std::thread([&, look, step]() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
          cv_a.wait(lk, [&] { return alloc_stop || failed || (populate_ok ? std::max(alloc_end, next_off) : alloc_end) < next_off + look; });
          if (alloc_stop || failed) return;
          const int64_t from = populate_ok ? std::max(alloc_end, next_off) : alloc_end;
          lk.unlock();
          int e = 0;
          while (fallocate(out_fd, FALLOC_FL_KEEP_SIZE, (off_t)from, (off_t)step) != 0) { if (errno != EINTR) { e = errno; break; } }
          lk.lock();
          if (e) {
            failed = true;
            errmsg = std::string("cannot reserve space for the SAM output: ") + strerror(e);
            wake_all();
            return;
          }
          alloc_end = from + step;
          if (!populate_ok && alloc_end > file_size) {   // writers wait for the allocator: the file has to cover what it reserved
            file_size = alloc_end;
            if (ftruncate(out_fd, (off_t)file_size) != 0) { failed = true; errmsg = "cannot extend the SAM output"; wake_all(); return; }
          }
          cv_w.notify_all();
        }
      });
    }
  }
  // the file is extended (no allocation) ahead of the offsets handed out; called with `mu` held
  auto grow_file = [&]() -> bool {
    if (!out_map || next_off <= file_size) return true;
    file_size = next_off + ((int64_t)1 << 30);
    return ftruncate(out_fd, (off_t)file_size) == 0;
  };
  std::deque<std::shared_ptr<Job>> wq;                      // jobs with an assigned offset, waiting to be written
  std::vector<char> busy((size_t)G * n_workers * 2, 0);    // pinned buffer [device][worker][slot] still being written out
  int gpu_active = G * n_workers;
  const bool times = getenv("B200BWA_TIMES") != 0;
  std::vector<std::thread> writers;
  if (positioned) {
    int n_writers = 3 * G;
    if (const char* s = getenv("B200BWA_WRITERS")) n_writers = atoi(s);
    if (n_writers < 1) n_writers = 1;
    for (int t = 0; t < n_writers; ++t)
      writers.emplace_back([&]() {
        for (;;) {
          std::shared_ptr<Job> j;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv_w.wait(lk, [&] { return !wq.empty() || failed || gpu_active == 0; });
            if (failed || wq.empty()) return;
            j = wq.front();
            wq.pop_front();
          }
          bool wrote = true;
          if (out_map && !populate_ok) {
            std::unique_lock<std::mutex> lk(mu);
            cv_w.wait(lk, [&] { return failed || alloc_end >= j->offset + (int64_t)j->sam_size; });
            if (failed) return;
          }
          const double t_w0 = b2_realtime();
          if (out_map) {
            if (populate_ok && j->sam_size) {
              const int64_t a0 = j->offset & ~(int64_t)4095, a1 = (j->offset + (int64_t)j->sam_size + 4095) & ~(int64_t)4095;
              if (madvise(out_map + a0, (size_t)(a1 - a0), MADV_POPULATE_WRITE) != 0) {
                std::lock_guard<std::mutex> lk(mu);
                failed = true;
                errmsg = std::string("cannot reserve space for the SAM output: ") + strerror(errno == EFAULT ? ENOSPC : errno);
                j->rc = 1;
                wake_all();
                return;
              }
            }
            memcpy(out_map + j->offset, j->wdata, j->sam_size);
            {  // drop this batch's page-table entries now, in this thread (the pages stay in the page cache: the mapping is
               // shared): tearing down the whole mapping's entries at the end of the call cost ~35 ns per KB of SAM, serially
              const int64_t d0 = (j->offset + 4095) & ~(int64_t)4095, d1 = (j->offset + (int64_t)j->sam_size) & ~(int64_t)4095;
              if (d1 > d0 && !getenv("B200BWA_KEEP_PTES")) madvise(out_map + d0, (size_t)(d1 - d0), MADV_DONTNEED);
            }
          } else wrote = pwrite_all(j->wdata, j->sam_size, j->offset);
          if (times) fprintf(stderr, "[T::b200bwa] batch %ld: queued for writing %.1f ms, write %.1f ms, done at %.1f ms\n", (long)j->index, (t_w0 - j->t_gpu) * 1e3, (b2_realtime() - t_w0) * 1e3, (b2_realtime() - t_real) * 1e3);
          std::lock_guard<std::mutex> lk(mu);
          if (j->wslot >= 0) { busy[j->wslot] = 0; cv_dev[j->wslot / (2 * n_workers)].notify_all(); }
          if (!wrote) { j->rc = 1; failed = true; errmsg = "write error on SAM output"; wake_all(); }
          j->done = true;
          cv_m.notify_all();
        }
      });
  }

  if (times0) fprintf(stderr, "[T::b200bwa] pipeline threads starting at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  // A job of a device goes to its lowest-numbered idle worker: a device that never has more than three batches in
  // flight keeps using the same three stream/buffer sets (their buffers are allocated once, and stay warm).
  std::vector<unsigned> idle(G, 0u);
  std::vector<std::thread> gpu;
  for (int dv = 0; dv < G; ++dv)
   for (int wi = 0; wi < n_workers; ++wi)
    gpu.emplace_back([&, dv, wi]() {
      B2HostBatch hb;
      B2Engine* eng = engs[dv];
      std::deque<std::shared_ptr<Job>>& q = todo[dv];
      struct Leave { std::mutex& m; std::condition_variable& c; int& n; ~Leave() { std::lock_guard<std::mutex> lk(m); --n; c.notify_all(); } } leave{mu, cv_w, gpu_active};
      for (;;) {
        std::shared_ptr<Job> j;
        int slot = -1;
        {
          std::unique_lock<std::mutex> lk(mu);
          idle[dv] |= 1u << wi;
          cv_dev[dv].wait(lk, [&] { return (!q.empty() && (idle[dv] & ((1u << wi) - 1)) == 0) || (q.empty() && read_done) || failed; });
          idle[dv] &= ~(1u << wi);
          if (failed || (q.empty() && read_done)) { cv_dev[dv].notify_all(); return; }
          j = q.front();
          q.pop_front();
          cv_dev[dv].notify_all();   // (a higher-numbered worker may be the lowest idle one now)
          if (positioned && !smart) {  // the pinned buffer this batch's text will land in must have been written out
            slot = (dv * n_workers + wi) * 2 + eng->next_view_slot(wi);
            cv_dev[dv].wait(lk, [&] { return !busy[slot] || failed; });
            if (failed) return;
          }
        }
        double ct = b2_cputime(), rt = b2_realtime();
        const char* data = 0;
        size_t len = 0;
        int rc;
        const bool raw_path = j->raw->compact && !smart;
        int n_reads_job = (int)j->raw->n_reads();
        if (raw_path) {
          // plain FASTQ from mapped files: ship the bytes, the device packs (k_ingest_*)
          B2RawInput in;
          const B2RawBatch& rb = *j->raw;
          in.n_files = rb.n_files; in.n_reads = n_reads_job; in.max_len = rb.max_len; in.copy_comment = m.copy_comment;
          in.n_bases = rb.n_bases; in.n_processed = rb.n_processed;
          for (int f = 0; f < rb.n_files; ++f) { in.base[f] = rb.cbase[f]; in.recs[f] = rb.crec[f].data(); in.n_recs[f] = rb.crec[f].size(); }
          rc = positioned ? eng->process_raw_view(in, pes0, wi, &data, &len) : eng->process_raw(in, pes0, j->sam, wi);
          j->raw.reset();
        } else {
          expand_compact(j->raw.get());
          b2_pack_batch(*j->raw, m.copy_comment, &hb);
          j->raw.reset();
        }
        const double t_packed = b2_realtime();
        if (raw_path) {
        } else if (smart) {
          rc = process_smart(eng, hb, pes0, wi, m.verbose, j->sam);
          data = j->sam.data(); len = j->sam.size();
        } else rc = positioned ? eng->process_view(hb, pes0, wi, &data, &len) : eng->process(hb, pes0, j->sam, wi);
        if (m.verbose >= 3 && !rc)
          fprintf(stderr, "[M::mem_process_seqs] Processed %d reads in %.3f CPU sec, %.3f real sec\n", n_reads_job,
                  b2_cputime() - ct, b2_realtime() - rt);
        const double t_gpu = b2_realtime();
        if (getenv("B200BWA_TIMES") && !rc) {
          const B2StageTimes& t = eng->times(wi);
          fprintf(stderr, "[T::b200bwa] device %d worker %d batch %ld (%s): start %.1f ms, pack %.1f ms, upload+kernels+download %.1f ms (wall)\n", eng->device(), wi,
                  (long)j->index, raw_path ? "device ingest" : "host pack", (rt - t_real) * 1e3, (t_packed - rt) * 1e3, (t_gpu - t_packed) * 1e3);
          fprintf(stderr, "[T::b200bwa] worker %d reads %d: stage %.2f h2d %.2f seed %.2f sa %.2f chain %.2f extend %.2f pestat %.2f rescue %.2f final %.2f gather %.2f d2h %.2f total %.2f ms; launches %d seeds %ld regs %ld sam %ld B\n",
                  wi, n_reads_job, t.stage_host, t.h2d, t.seed, t.sa, t.chain, t.extend, t.pestat, t.rescue, t.final_, t.gather, t.d2h, t.total,
                  t.launches, (long)t.n_seed_tasks, (long)t.n_regs, (long)t.sam_bytes);
        }
        if (positioned && !rc) {
          std::lock_guard<std::mutex> lk(mu);
          j->sam_size = len; j->size_known = true; j->wdata = data; j->wslot = slot; j->t_gpu = t_gpu;
          if (slot >= 0) busy[slot] = 1;
          while (!unassigned.empty() && unassigned.front()->size_known) {  // hand out offsets along the known prefix
            unassigned.front()->offset = next_off;
            next_off += (int64_t)unassigned.front()->sam_size;
            wq.push_back(unassigned.front());
            unassigned.pop_front();
          }
          if (!grow_file()) { failed = true; errmsg = "cannot extend the SAM output"; wake_all(); }
          cv_w.notify_all(); cv_a.notify_all();
          continue;   // a writer thread marks the job done
        }
        std::lock_guard<std::mutex> lk(mu);
        j->rc = rc; j->done = true;
        if (rc) { failed = true; wake_all(); }
        cv_m.notify_all();
      }
    });

  int rc = write_failed ? 1 : 0;
  if (write_failed) errmsg = "write error on SAM output";
  for (;;) {
    std::shared_ptr<Job> j;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_m.wait(lk, [&] { return failed || (!inorder.empty() && inorder.front()->done) || (inorder.empty() && read_done); });
      if (failed) { rc = 1; break; }
      if (inorder.empty()) break;
      j = inorder.front();
      inorder.pop_front();
      cv_r.notify_all();
    }
    if (fidx) fprintf(fidx, "%ld\t%d\t%zu\n", (long)j->index, j->n_reads, positioned ? j->sam_size : j->sam.size());
    if (!positioned && fwrite(j->sam.data(), 1, j->sam.size(), out) != j->sam.size()) {
      std::lock_guard<std::mutex> lk(mu);
      failed = true; rc = 1; errmsg = "write error on SAM output";
      wake_all();
      break;
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (rc) failed = true;
    wake_all();
  }
  if (times0) fprintf(stderr, "[T::b200bwa] last batch delivered at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  reader.join();
  for (auto& t : gpu) t.join();
  for (auto& t : writers) t.join();
  if (allocator.joinable()) {
    { std::lock_guard<std::mutex> lk(mu); alloc_stop = true; wake_all(); }
    allocator.join();
  }
  if (times0) fprintf(stderr, "[T::b200bwa] threads joined at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  if (out_map) {
    munmap(out_map, out_map_len);
    if (times0) fprintf(stderr, "[T::b200bwa] output unmapped at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
    if (!rc && ftruncate(out_fd, (off_t)next_off) != 0) { rc = 1; errmsg = "cannot set the final size of the SAM output"; }
    if (times0) fprintf(stderr, "[T::b200bwa] output truncated to size at %.1f ms\n", (b2_realtime() - t_real) * 1e3);
  }
  // a failed run leaves the header alone: the batches behind it were written at their final offsets as they finished, so
  // the reserved range has holes (and, mapped, its fallocate()d length)
  if (rc && positioned && ftruncate(out_fd, (off_t)hdr_bytes) != 0) {}
  if (!rc) {   // the reference closes its inputs after the last batch is out and exits with 1 if that fails: output complete, status 1
    ks.close();
    if (have2) ks2.close();
    const std::string& e = !ks.io_error().empty() || !have2 ? ks.io_error() : ks2.io_error();
    if (!e.empty()) { rc = 1; errmsg = e; }
  }
  if (rc && m.verbose >= 1 && !errmsg.empty() && errmsg[0] == '[') fprintf(stderr, "%s\n", errmsg.c_str());
  if (rc) {
    std::string em = errmsg;
    for (B2Engine* e : engs) if (em.empty() && e->last_error()[0]) em = e->last_error();
    set_err("%s", em.c_str());
  }
  if (fidx) fclose(fidx);
  if (positioned) { if (::close(out_fd) != 0 && !rc) { rc = 1; set_err("close error on SAM output"); } }
  else if (out_path) { if (fclose(out) != 0 && !rc) { rc = 1; set_err("close error on SAM output"); } }
  else fflush(out);
  if (rc == 0 && m.verbose >= 3) {
    fprintf(stderr, "[main] Version: %s\n", B200BWA_VERSION);
    fprintf(stderr, "[main] CMD:");
    if (pg_cmdline) { const char* p = strstr(pg_cmdline, "CL:"); fprintf(stderr, " %s", p ? p + 3 : ""); }
    fprintf(stderr, "\n[main] Real time: %.3f sec; CPU: %.3f sec\n", b2_realtime() - t_real, b2_cputime());
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------
// The step after the path: SparkBWA's reducer (BwaInterpreter.java:342-376) concatenates the SAM
// files of the partitions in partition order, keeping lines that start with '@' only from the first file, every line
// terminated by '\n', and deletes the inputs.
int b2_reduce_sam(int n_inputs, const char* const* inputs, const char* out_path, int delete_inputs) {
  FILE* out = fopen(out_path, "wb");
  if (!out) return set_err("fail to open output `%s'", out_path);
  setvbuf(out, 0, _IOFBF, 1 << 22);
  std::vector<char> buf((size_t)1 << 22);
  for (int i = 0; i < n_inputs; ++i) {
    FILE* in = fopen(inputs[i], "rb");
    if (!in) { fclose(out); return set_err("fail to open file `%s'", inputs[i]); }
    bool at_line_start = true, keep = true, any = false;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), in)) > 0) {
      size_t p = 0;
      while (p < got) {
        if (at_line_start) { keep = i == 0 || buf[p] != '@'; at_line_start = false; }
        const char* nl = (const char*)memchr(buf.data() + p, '\n', got - p);
        const size_t e = nl ? (size_t)(nl - buf.data()) + 1 : got;
        if (keep && fwrite(buf.data() + p, 1, e - p, out) != e - p) { fclose(in); fclose(out); return set_err("write error on `%s'", out_path); }
        any = true;
        if (nl) at_line_start = true;
        p = e;
      }
    }
    if (any && !at_line_start && keep) fputc('\n', out);  // readLine() + "\n": an unterminated last line gets its newline
    fclose(in);
  }
  if (fclose(out) != 0) return set_err("close error on `%s'", out_path);
  if (delete_inputs) for (int i = 0; i < n_inputs; ++i) unlink(inputs[i]);
  return 0;
}

// Spark-free driver of the same shape as SparkBWA's map step (BwaInterpreter.java:301-319 over
// BwaPairedAlignment/BwaSingleAlignment): partition p of the reads -> one `mem` call -> <out_dir>/<app>-<p>.sam
// (BwaAlignmentBase.java:176-178), then optionally the reducer into <out_dir>/FullOutput.sam.
int b2_run_partitions(int n_opts, char** opts, const char* idxbase, int n_parts, const char* const* fq1, const char* const* fq2,
                      const char* out_dir, const char* app, int reduce) {
  std::vector<std::string> outs;
  for (int p = 0; p < n_parts; ++p) {
    char num[32];
    snprintf(num, sizeof(num), "-%d.sam", p);
    outs.push_back(std::string(out_dir) + "/" + (app ? app : "SparkBWA") + num);
    std::vector<std::string> a;
    a.push_back("mem");
    for (int i = 0; i < n_opts; ++i) a.push_back(opts[i]);
    a.push_back(idxbase);
    a.push_back(fq1[p]);
    if (fq2 && fq2[p]) a.push_back(fq2[p]);
    std::vector<char*> argv;
    for (auto& x : a) argv.push_back(&x[0]);
    std::string pg = "@PG\tID:bwa\tPN:bwa\tVN:" B200BWA_VERSION "\tCL:bwa";
    for (auto& x : a) { pg += ' '; pg += x; }
    if (b2_mem_main((int)argv.size(), argv.data(), outs.back().c_str(), pg.c_str())) return 1;
  }
  if (reduce) {
    std::vector<const char*> in;
    for (auto& o : outs) in.push_back(o.c_str());
    return b2_reduce_sam((int)in.size(), in.data(), (std::string(out_dir) + "/FullOutput.sam").c_str(), 1);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// `bwa index` twin: bwa_index / bwa_idx_build (bwa/bwtindex.c:210-324) with the BWT, Occ and SA
// built on the GPU (b2_index.cu).  .pac/.ann/.amb follow bns_fasta2bntseq + bns_dump (bwa/bntseq.c:66-96,228-328):
// ambiguous bases become lrand48() & 3 with srand48(11) -- replayed here with the same 48-bit LCG so the packed
// sequence, and with it every index file, is byte-identical -- and each run of one ambiguous letter is a "hole".
int b2_build_fm(const uint8_t* fwd, int64_t l_pac, const char* prefix, int device, int64_t group_max, int sa_intv, int verbose,
                std::string* err);

int b2_index_main(int argc, char** argv) {
  std::string prefix, fa;
  bool is_64 = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-6") is_64 = true;
    else if ((a == "-p" || a == "-a" || a == "-b") && i + 1 < argc) {
      ++i;
      if (a == "-p") prefix = argv[i];
      else if (a == "-a" && strcmp(argv[i], "rb2") && strcmp(argv[i], "bwtsw") && strcmp(argv[i], "is"))
        return set_err("unknown algorithm: '%s'.", argv[i]);   // (the three algorithms give the same files; so does this one)
    } else if (a.size() > 1 && a[0] == '-') return set_err("index: invalid option '%s'", a.c_str());
    else if (fa.empty()) fa = a;
  }
  if (fa.empty()) {
    fprintf(stderr, "\nUsage:   bwa index [-p prefix] [-6] <in.fasta>\n\n");
    return set_err("index: no input file");
  }
  if (prefix.empty()) prefix = fa + (is_64 ? ".64" : "");
  const int verbose = 3;
  double t0 = b2_realtime();
  B2SeqReader ks;
  if (ks.open(fa.c_str())) { fprintf(stderr, "[E::bwa_idx_build] fail to open file '%s'\n", fa.c_str()); return set_err("fail to open file '%s'", fa.c_str()); }
  struct Ann { std::string name, anno; int64_t offset; int len, n_ambs; };
  struct Amb { int64_t offset; int len; char amb; };
  std::vector<Ann> anns;
  std::vector<Amb> ambs;
  std::vector<uint8_t> fwd;
  uint64_t lcg = ((uint64_t)11 << 16) | 0x330E;   // srand48(11)
  const unsigned char* nt4 = b2_nt4_table();
  while (ks.next() >= 0) {
    Ann p;
    p.name = ks.name();
    p.anno = ks.comment().empty() ? "(null)" : ks.comment();
    p.len = (int)ks.seq().size();
    p.offset = anns.empty() ? 0 : anns.back().offset + anns.back().len;
    p.n_ambs = 0;
    const std::string& sq = ks.seq();
    const size_t base = fwd.size();
    fwd.resize(base + sq.size());
    uint8_t* dst = fwd.data() + base;
    unsigned any_amb = 0;
    for (size_t i = 0; i < sq.size(); ++i) { const uint8_t c = nt4[(unsigned char)sq[i]]; dst[i] = c; any_amb |= c; }
    if (any_amb & 4) {  // ambiguous bases present: holes + the lrand48 replay, in sequence order
      int lasts = 0;
      for (size_t i = 0; i < sq.size(); ++i) {
        if (dst[i] >= 4) {
          if (lasts == sq[i]) ++ambs.back().len;
          else { Amb q; q.len = 1; q.offset = p.offset + (int64_t)i; q.amb = sq[i]; ambs.push_back(q); ++p.n_ambs; }
          lcg = (lcg * 0x5DEECE66DULL + 0xB) & 0xFFFFFFFFFFFFULL;   // lrand48()
          dst[i] = (uint8_t)((lcg >> 17) & 3);
        }
        lasts = sq[i];
      }
    }
    anns.push_back(p);
  }
  ks.close();
  const int64_t l_pac = (int64_t)fwd.size();
  if (l_pac == 0) return set_err("index: no sequence in '%s'", fa.c_str());
  if (verbose >= 3) fprintf(stderr, "[bwa_index] Pack FASTA... %.2f sec\n", b2_realtime() - t0);
  {  // .pac, forward strand only (the final state after bwa_idx_build's second packing pass)
    std::vector<uint8_t> pac((size_t)(l_pac >> 2) + ((l_pac & 3) ? 1 : 0), 0);
    for (int64_t l = 0; l < l_pac; ++l) pac[l >> 2] |= fwd[l] << ((~l & 3) << 1);
    FILE* fp = fopen((prefix + ".pac").c_str(), "wb");
    if (!fp) return set_err("cannot write %s.pac", prefix.c_str());
    fwrite(pac.data(), 1, pac.size(), fp);
    if (l_pac % 4 == 0) fputc(0, fp);
    fputc((int)(l_pac % 4), fp);
    if (fclose(fp) != 0) return set_err("write error on %s.pac", prefix.c_str());
  }
  {
    FILE* fp = fopen((prefix + ".ann").c_str(), "w");
    if (!fp) return set_err("cannot write %s.ann", prefix.c_str());
    fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)anns.size(), 11u);
    for (const Ann& p : anns) {
      fprintf(fp, "%d %s", 0, p.name.c_str());
      if (!p.anno.empty()) fprintf(fp, " %s\n", p.anno.c_str());
      else fprintf(fp, "\n");
      fprintf(fp, "%lld %d %d\n", (long long)p.offset, p.len, p.n_ambs);
    }
    if (fclose(fp) != 0) return set_err("write error on %s.ann", prefix.c_str());
    fp = fopen((prefix + ".amb").c_str(), "w");
    if (!fp) return set_err("cannot write %s.amb", prefix.c_str());
    fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)anns.size(), (unsigned)ambs.size());
    for (const Amb& q : ambs) fprintf(fp, "%lld %d %c\n", (long long)q.offset, q.len, q.amb);
    if (fclose(fp) != 0) return set_err("write error on %s.amb", prefix.c_str());
  }
  t0 = b2_realtime();
  if (verbose >= 3) fprintf(stderr, "[bwa_index] Construct BWT, Occ and SA for the packed sequence...\n");
  int device = getenv("B200BWA_DEVICE") ? atoi(getenv("B200BWA_DEVICE")) : 0;
  int64_t group_max = getenv("B200BWA_INDEX_GROUP") ? atoll(getenv("B200BWA_INDEX_GROUP")) : 0;
  std::string err;
  if (b2_build_fm(fwd.data(), l_pac, prefix.c_str(), device, group_max, 32, verbose, &err)) {
    fprintf(stderr, "[E::b200bwa_index] %s\n", err.c_str());
    return set_err("%s", err.c_str());
  }
  if (verbose >= 3) fprintf(stderr, "[bwa_index] %.2f seconds elapse.\n", b2_realtime() - t0);
  return 0;
}
