// Launch/memory shim.  The product build (nvcc, sm_100a) maps these onto the CUDA runtime.
// Defining B2_HOSTSIM (done ONLY by tests/hostsim, never by the shipped library) turns a launch into
// a serial loop over the same kernel bodies so the control logic can be checked on a machine
// without a GPU; it is test scaffolding, not a fallback, and libbwa.so does not contain it.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(B2_HOSTSIM)
struct B2Dim { int grid, block, tid; };
extern thread_local B2Dim b2_dim;
#define B2_KERNEL static void
#define B2_KERNEL_LB(threads, blocks_per_sm) static void
#define B2_GTID() (b2_dim.tid)
#define B2_GSIZE() (b2_dim.grid * b2_dim.block)
#define B2_LAUNCH(kern, _g, _b, stream, ...)                             \
  do {                                                                   \
    b2_dim.grid = (_g); b2_dim.block = (_b);                             \
    for (int _t = 0; _t < (_g) * (_b); ++_t) { b2_dim.tid = _t; kern(__VA_ARGS__); } \
  } while (0)
#define B2_LAUNCH_SMEM(kern, _g, _b, _smem, stream, ...) B2_LAUNCH(kern, _g, _b, stream, __VA_ARGS__)
typedef int b2_stream_t;
inline int b2_dev_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
inline void b2_dev_free(void* p) { free(p); }
inline int b2_h2d(void* d, const void* h, size_t n, b2_stream_t) { memcpy(d, h, n); return 0; }
inline int b2_d2h(void* h, const void* d, size_t n, b2_stream_t) { memcpy(h, d, n); return 0; }
inline int b2_dev_memset(void* d, int v, size_t n, b2_stream_t) { memset(d, v, n); return 0; }
inline int b2_sync(b2_stream_t) { return 0; }
inline int b2_host_alloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
inline void b2_host_free(void* p) { free(p); }
#define B2_ATOMIC_OR(p, v) (*(p) |= (v))
template <class T, class V>
inline T b2_host_fetch_add(T* p, V v) { T o = *p; *p += (T)v; return o; }
#define B2_ATOMIC_ADD(p, v) b2_host_fetch_add((p), (v))
#define B2_ATOMIC_CAS(p, cmp, val) (*(p) == (cmp) ? (*(p) = (val), (cmp)) : *(p))
#define B2_ATOMIC_MIN(p, v) (*(p) = *(p) < (v) ? *(p) : (v))
#define B2_WARP_LANES 1
#define B2_TASK_GROUP 1
#define B2_G_EXT 1
#define B2_G_FINAL 1
#define B2_G_HEAVY 1
#define B2_G_RESC 1
#define B2_LB_EXT 1
#define B2_LB_FINAL 1
#else
#include <cuda_runtime.h>
#define B2_KERNEL static __global__ void
#define B2_KERNEL_LB(threads, blocks_per_sm) static __global__ void __launch_bounds__(threads, blocks_per_sm)
#define B2_GTID() ((int)(blockIdx.x * blockDim.x + threadIdx.x))
#define B2_GSIZE() ((int)(gridDim.x * blockDim.x))
// B200BWA_SYNC_DEBUG=1: wait for every kernel and name the one that faults (diagnosis only: serialises the stream)
inline void b2_launch_check(const char* name, cudaStream_t s) {
  static const int dbg = getenv("B200BWA_SYNC_DEBUG") ? atoi(getenv("B200BWA_SYNC_DEBUG")) : 0;
  if (!dbg) return;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) fprintf(stderr, "[E::b200bwa] kernel %s: %s\n", name, cudaGetErrorString(e));
}
#define B2_LAUNCH(kern, _g, _b, stream, ...) do { kern<<<(_g), (_b), 0, (stream)>>>(__VA_ARGS__); b2_launch_check(#kern, (stream)); } while (0)
#define B2_LAUNCH_SMEM(kern, _g, _b, _smem, stream, ...) do { kern<<<(_g), (_b), (_smem), (stream)>>>(__VA_ARGS__); b2_launch_check(#kern, (stream)); } while (0)
typedef cudaStream_t b2_stream_t;
inline int b2_dev_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) != cudaSuccess; }
inline void b2_dev_free(void* p) { if (p) cudaFree(p); }
inline int b2_h2d(void* d, const void* h, size_t n, b2_stream_t s) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) != cudaSuccess; }
inline int b2_d2h(void* h, const void* d, size_t n, b2_stream_t s) { return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) != cudaSuccess; }
inline int b2_dev_memset(void* d, int v, size_t n, b2_stream_t s) { return cudaMemsetAsync(d, v, n, s) != cudaSuccess; }
inline int b2_sync(b2_stream_t s) { return cudaStreamSynchronize(s) != cudaSuccess; }
inline int b2_host_alloc(void** p, size_t n) { return cudaMallocHost(p, n ? n : 1) != cudaSuccess; }
inline void b2_host_free(void* p) { if (p) cudaFreeHost(p); }
#define B2_ATOMIC_OR(p, v) atomicOr((p), (v))
#define B2_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define B2_ATOMIC_CAS(p, cmp, val) atomicCAS((p), (cmp), (val))
#define B2_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define B2_WARP_LANES 32
#define B2_TASK_GROUP B2_TASK_LANES
// Lanes per task of the group-cooperative kernels.  The chain-extension kernel runs the register-resident row-parallel
// ksw_extend2 on whole warps (constant member mask); the rescue SW keeps the 16 lanes of ksw_u8, whose two groups per
// warp run the same loop nearly in lock step; the finalisation kernel, whose DP now runs ahead of it (b2_galn.cuh),
// is serial bookkeeping plus bulk copies, and does best with small groups: more pairs in flight per register.
#ifndef B2_G_EXT
#define B2_G_EXT 32
#endif
#ifndef B2_G_FINAL
#define B2_G_FINAL 4   // measured on the B200 (profiles/r1d/variants_group_size.jsonl): 32: 5.20, 16: 5.38, 8: 5.69, 4: 5.83, 2: 5.34, 1: 5.36 M reads/s
#endif
#define B2_G_RESC 16
#ifndef B2_G_HEAVY
#define B2_G_HEAVY 32  // pairs with many regions (k_final_pe_heavy): a warp each
#endif
#ifndef B2_LB_EXT
#define B2_LB_EXT 4
#endif
#ifndef B2_LB_FINAL
#define B2_LB_FINAL 3
#endif
#endif
