// INT32 / DPX issue-rate micro-benchmark for the SW rooflines ("INT peak to be measured by a
// micro-benchmark on the box").  Each thread runs 8 independent dependency chains of one instruction kind; the result
// is lane-operations per second over the whole chip.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu && ./int_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

template <int OP>
__device__ __forceinline__ int step(int a, int b, int c) {
  int r;
  // (plain C for these four would be folded across the unrolled chain: volatile asm keeps every instruction)
  if (OP == 0) { asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }                    // IADD3
  if (OP == 5) { asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }     // IMAD
  if (OP == 6) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; } // LOP3
  if (OP == 8) { asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }                    // IMNMX
  if (OP == 1) return __viaddmax_s32(a, b, c);                 // VIADDMNMX: max(a + b, c)
  if (OP == 2) return __vimax3_s32(a, b, c);                   // VIMNMX3
  if (OP == 3) return (int)__viaddmax_s16x2((unsigned)a, (unsigned)b, (unsigned)c);   // packed 2 x s16 add+max
  if (OP == 4) return (int)__vimax3_s16x2((unsigned)a, (unsigned)b, (unsigned)c);     // packed 2 x s16 max3
  if (OP == 7) return __popc(a) + b;                           // POPC (+ IADD)
  if (OP == 9) return (int)__viaddmax_s16x2_relu((unsigned)a, (unsigned)b, (unsigned)c);
  return a;
}

template <int OP>
__global__ void __launch_bounds__(256) k_ops(int* out, int iters, int seed) {
  int v[CHAINS];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) v[k] = seed * (k + 1) + t;
  const int b = seed | 1, c = seed ^ 0x55;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) v[k] = step<OP>(v[k], b, c + u);
    }
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s ^= v[k];
  if (s == 0x7fffffff) out[t] = s;
}

template <int OP>
double run(int* d_out, int grid, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_ops<OP><<<grid, 256>>>(d_out, iters / 8, 3);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_ops<OP><<<grid, 256>>>(d_out, iters, 3 + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)grid * 256 * (double)iters * UNROLL * CHAINS;
  return ops / (best * 1e-3) / 1e12;  // tera lane-ops / s
}

int main() {
  int dev = 0, sms = 0, clk = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  int* d_out;
  cudaMalloc(&d_out, (size_t)sms * 8 * 256 * sizeof(int));
  const int grid = sms * 8, iters = 4096;
  const char* names[] = {"iadd3", "viaddmax_s32", "vimax3_s32", "viaddmax_s16x2", "vimax3_s16x2", "imad", "lop3", "popc_add", "imnmx", "viaddmax_s16x2_relu"};
  double r[10];
  r[0] = run<0>(d_out, grid, iters); r[1] = run<1>(d_out, grid, iters); r[2] = run<2>(d_out, grid, iters);
  r[3] = run<3>(d_out, grid, iters); r[4] = run<4>(d_out, grid, iters); r[5] = run<5>(d_out, grid, iters);
  r[6] = run<6>(d_out, grid, iters); r[7] = run<7>(d_out, grid, iters); r[8] = run<8>(d_out, grid, iters);
  r[9] = run<9>(d_out, grid, iters);
  if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "CUDA error\n"); return 1; }
  printf("{\"sms\": %d, \"clock_khz_max\": %d, \"unit\": \"tera lane-instructions/s (whole chip, 8 independent chains per thread, 2048 threads/SM)\"", sms, clk);
  for (int i = 0; i < 10; ++i) printf(", \"%s\": %.3f", names[i], r[i]);
  printf(", \"note\": \"popc_add issues two instructions per step (POPC + IADD); packed s16x2 forms process two 16-bit cells per lane-instruction\"}\n");
  return 0;
}
