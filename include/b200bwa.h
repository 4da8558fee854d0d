/* b200bwa -- C-ABI of the B200-native BWA-MEM engine (drop-in for the `mem` path of SparkBWA's libbwa.so).
 *
 * Every entry point uses plain C types only.  Reference interfaces replaced (paths relative to
 * citiususc/SparkBWA, `bwa/` = src/main/native/bwa-0.7.15/):
 *
 *   Java_com_github_sparkbwa_BwaJni_bwa_1jni   src/main/native/bwa_jni.c:29, declared in
 *                                              src/main/native/com_github_sparkbwa_BwaJni.h:15-16;
 *                                              bound by java/com/github/sparkbwa/BwaJni.java:59
 *   b200bwa_main                               bwa/main.c:61 (`int main(int argc, char *argv[])`), `mem` only
 *   b200bwa_mem                                bwa/fastmap.c:115 (`int main_mem(int argc, char *argv[])`)
 *   b200bwa_index_build                        bwa/bwtindex.c:210 (`int bwa_index(int argc, char *argv[])`)
 *   b200bwa_index_load / _align_batch / ...    bwa/bwa.c:252 (bwa_idx_load_from_disk) and
 *                                              bwa/bwamem.c:1172 (mem_process_seqs) -- the library-level
 *                                              calls a host program (bwa/example.c) would make
 *
 * Error behaviour: nothing in this library calls exit()/abort() (the reference's err_fatal does,
 * bwa/utils.c:90-122, and would take an executor JVM down); failures return non-zero and the
 * message is available from b200bwa_last_error().  There is no CPU fallback: without a usable
 * CUDA device every aligning call fails.
 */
#ifndef B200BWA_H
#define B200BWA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* argv-compatible with `bwa` (argv[0] = program name, argv[1] = "mem", then bwa-mem options,
 * index prefix, reads [, mates]).  SAM goes to stdout.  Other sub-commands return 1.
 * One call aligns ONE job on every visible GPU (B200BWA_DEVICES="all" | "0,2,5", or B200BWA_DEVICE=n for one):
 * the -K batches are dealt round-robin over the devices and their SAM blocks are written in input order, as
 * kt_pipeline orders them (bwa/kthread.c:92-102, bwa/fastmap.c:84-89). */
int b200bwa_main(int argc, char** argv);

/* Same, but the SAM text is written to `out_path` (what the JNI glue achieves with
 * "-f <out>" + freopen, bwa_jni.c:74-135) without touching the process-wide stdout. */
int b200bwa_main_to(int argc, char** argv, const char* out_path);

/* argv as main_mem sees it: argv[0] = "mem". out_path may be NULL (stdout). */
int b200bwa_mem(int argc, char** argv, const char* out_path);

/* `bwa index` (bwa/bwtindex.c:210-324 bwa_index / bwa_idx_build; bwa/bntseq.c:275-328 bns_fasta2bntseq;
 * bwa/bwt.c:62-84 bwt_cal_sa): argv[0] = "index", then [-p prefix] [-6] [-a algo] <in.fasta>.  Writes
 * <prefix>.pac/.ann/.amb/.bwt/.sa byte-identical to the reference's, the suffix sorting done on the GPU
 * (B200BWA_DEVICE selects it).  Also reachable as `b200bwa_main({"bwa", "index", ...})`. */
int b200bwa_index_build(int argc, char** argv);

/* The file-driven entry points keep the device image of an index cached across calls (keyed by prefix and the
 * size/mtime of <prefix>.bwt), the way `bwa shm` keeps it in host shared memory (bwa/bwashm.c:87-118).  This
 * releases every cached image that no running call uses; returns the number released. */
int b200bwa_index_cache_evict(void);

/* The steps either side of the path, Spark-free:
 * b200bwa_reduce_sam     SparkBWA's reducer, java/com/github/sparkbwa/BwaInterpreter.java:342-376: concatenates the
 *                        partitions' SAM files in order, keeps '@' lines of the first file only, deletes the inputs
 *                        when delete_inputs != 0 (the reference always does).
 * b200bwa_run_partitions the map step, BwaInterpreter.java:301-319 over BwaPairedAlignment / BwaSingleAlignment:
 *                        partition p (fq1[p] [, fq2[p]]; fq2 may be NULL) -> one `mem` call with the option tokens
 *                        `opts` -> <out_dir>/<app_name>-<p>.sam (BwaAlignmentBase.java:176-178); reduce != 0 then merges
 *                        them into <out_dir>/FullOutput.sam. */
int b200bwa_reduce_sam(int n_inputs, const char* const* inputs, const char* out_path, int delete_inputs);
int b200bwa_run_partitions(int n_opts, char** opts, const char* idxbase, int n_parts, const char* const* fq1,
                           const char* const* fq2, const char* out_dir, const char* app_name, int reduce);

/* Last error message of the calling thread's most recent failing call ("" if none). */
const char* b200bwa_last_error(void);

/* ---- library-level API (device-resident index, batches in host memory) -------------------- */

typedef struct b200bwa_index b200bwa_index_t; /* opaque: FM-index image resident in HBM */

/* Loads <prefix>.bwt/.sa/.pac/.ann/.amb[/.alt] and uploads them to `device`.  NULL on failure. */
b200bwa_index_t* b200bwa_index_load(const char* prefix, int device);
void b200bwa_index_free(b200bwa_index_t* idx);

/* Sets bwa-mem options from an option string as given to `-w` of SparkBWA / the bwa command line
 * (e.g. "-t 4 -K 10000000 -M"); NULL or "" = defaults.  Returns 0 on success.  Smart pairing (-p, bwa/fastmap.c:64-83)
 * is a property of the file-driven entry points above; b200bwa_align_batch takes its pairing from `paired`. */
int b200bwa_set_options(b200bwa_index_t* idx, const char* mem_options);

/* Aligns one batch held in HOST memory and returns the SAM records (no header) in a malloc'ed
 * buffer the caller frees with b200bwa_free.
 *   n_reads          number of reads; paired-end batches are interleaved (mate1, mate2, ...) and even
 *   seq/qual         concatenated bases (ASCII) / qualities; qual may be NULL (FASTA)
 *   seq_off[n+1]     start of read i in seq/qual (seq_off[n] = total)
 *   names            NUL-separated read names, one per read (the /1 /2 suffix already trimmed)
 *   n_processed      absolute index of the first read of this batch within the whole run
 *   paired           non-zero: paired-end; the two reads of a pair must then carry the same name, as in the reference
 *                    (bwa/bwamem_pair.c:355,385: "paired reads have different names" ends its run) -- rc 1 otherwise
 * Returns 0 on success. */
int b200bwa_align_batch(b200bwa_index_t* idx, int n_reads, const char* seq, const char* qual,
                        const int64_t* seq_off, const char* names, int64_t n_processed, int paired,
                        char** sam_out, size_t* sam_len);
void b200bwa_free(void* p);

/* Writes the SAM header (@SQ lines, -R/-H lines, @PG) for this index into a malloc'ed buffer. */
int b200bwa_sam_header(b200bwa_index_t* idx, const char* cmdline, char** out, size_t* out_len);

/* Device timing of the last b200bwa_align_batch on this index, milliseconds (CUDA events). */
typedef struct {
  float h2d, seed, sa, chain, extend, pestat, finalize, gather, d2h, total;
  int launches;
  int64_t n_seed_tasks, n_regs, sam_bytes;
  float rescue;   /* planned mate-rescue SW tasks (between pestat and finalize) */
  int64_t n_rescue_tasks;
  float k_xext, k_resc_sw, k_galn;  /* the DP kernels proper: inter-task ksw_extend2 (both sides), ksw_u8/i16, ksw_global2 */
} b200bwa_times_t;
int b200bwa_last_times(b200bwa_index_t* idx, b200bwa_times_t* t);

/* Benchmark helpers: keep a batch resident in HBM and run only the device pipeline. */
int b200bwa_upload_batch(b200bwa_index_t* idx, int n_reads, const char* seq, const char* qual,
                         const int64_t* seq_off, const char* names, int64_t n_processed);
int b200bwa_run_resident(b200bwa_index_t* idx, int paired, int want_sam, char** sam_out, size_t* sam_len);

/* Runs reads [first, first+count) of the batch uploaded by b200bwa_upload_batch on stream/buffer set
 * `worker` (0 .. b200bwa_n_workers()-1, else rc 1; sets may run concurrently from different host threads).  `paired` non-zero:
 * the range is paired-end (interleaved mates, even count).  `t` (may be NULL) receives the device stage times of this call. */
int b200bwa_run_resident_range(b200bwa_index_t* idx, int worker, int first, int count, int paired, int want_sam,
                               char** sam_out, size_t* sam_len, b200bwa_times_t* t);
int b200bwa_set_paired(b200bwa_index_t* idx, int paired);
/* Number of independent stream/buffer sets of this index's engine (B200BWA_WORKERS, default 6, at most 8). */
int b200bwa_n_workers(b200bwa_index_t* idx);

/* Algorithmic-work counters accumulated by the kernels of one worker: [0] 64-byte occ-block touches of
 * SMEM seeding, [1] SA lookups, [2] SA walk steps, [3] ksw_extend2 cells, [4] calls, [5] ksw_global2
 * cells, [6] calls, [7] local-SW cells, [8] calls (definitions: SURVEY.md 8(d)); [9] [10] [11] the cells of the kernels
 * k_xext_run, k_resc_sw and k_galn* alone (subsets of [3], [7], [5]). */
int b200bwa_counters(b200bwa_index_t* idx, int worker, unsigned long long out[16], int reset);
/* Process-wide totals: [0] host->device bytes, [1] device->host bytes, [2] kernel launches. */
int b200bwa_process_counters(unsigned long long out[3], int reset);

/* Multi-GPU plumbing: device pointers/sizes of the index image, and adoption of an image that a
 * peer already placed in this GPU's memory (e.g. by ncclBroadcast). */
int b200bwa_index_device_image(b200bwa_index_t* idx, void** bwt, size_t* bwt_bytes, void** sa, size_t* sa_bytes,
                               void** pac, size_t* pac_bytes);
b200bwa_index_t* b200bwa_index_adopt(const char* prefix, int device, const void* d_bwt, const void* d_sa,
                                     const void* d_pac);

#ifdef __cplusplus
}
#endif
#endif
