// TEST SCAFFOLDING ONLY: a single-threaded emulation of a CUDA lane group for the group-cooperative DP
// routines (sparkbwa_b200/csrc/b2_coopreg.cuh).  Every lane is a ucontext coroutine; a collective
// (shuffle, reduction, group barrier) deposits the lane's value and yields round-robin, so when control
// returns every lane of the group has reached the same collective -- the lock-step semantics the device
// code relies on.  Lets the cooperative kernels be fuzzed against the scalar forms on a GPU-less machine.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <ucontext.h>
#include <functional>

struct EmuWarp {
  enum { MAXL = 32, STACK = 256 * 1024 };
  int size = 0, cur = 0;
  ucontext_t main_ctx, ctx[MAXL];
  char* stacks[MAXL];
  bool done[MAXL];
  long long buf[2][MAXL];
  int phase[MAXL];
  std::function<void(int)> body;
  EmuWarp() { for (auto& s : stacks) s = (char*)malloc(STACK); }
  ~EmuWarp() { for (auto& s : stacks) free(s); }
  static void tramp(unsigned lo, unsigned hi) {
    EmuWarp* w = (EmuWarp*)(((uintptr_t)hi << 32) | lo);
    const int lane = w->cur;
    w->body(lane);
    w->done[lane] = true;
    w->yield();
  }
  void yield() {
    const int from = cur;
    for (int k = 1; k <= size; ++k) {
      int nx = (from + k) % size;
      if (!done[nx]) {
        if (nx == from) return;
        cur = nx;
        swapcontext(&ctx[from], &ctx[nx]);
        return;
      }
    }
    swapcontext(&ctx[from], &main_ctx);  // every lane finished
  }
  void run(int n, std::function<void(int)> f) {
    size = n; body = f;
    for (int i = 0; i < n; ++i) {
      done[i] = false; phase[i] = 0;
      getcontext(&ctx[i]);
      ctx[i].uc_stack.ss_sp = stacks[i];
      ctx[i].uc_stack.ss_size = STACK;
      ctx[i].uc_link = 0;
      uintptr_t p = (uintptr_t)this;
      makecontext(&ctx[i], (void (*)())tramp, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
    }
    cur = 0;
    swapcontext(&main_ctx, &ctx[0]);
  }
  long long xchg(int lane, long long v, int src) {
    const int ph = phase[lane]++ & 1;
    buf[ph][lane] = v;
    yield();
    return (src >= 0 && src < size) ? buf[ph][src] : v;
  }
  // all lanes' values of one collective (valid until the lane's next-but-one collective)
  const long long* gather(int lane, long long v) {
    const int ph = phase[lane]++ & 1;
    buf[ph][lane] = v;
    yield();
    return buf[ph];
  }
};

// Same member set as the device B2G (b2_coop.cuh)
struct EmuG {
  EmuWarp* w;
  unsigned mask;
  int lane, size;
  int shfl(int v, int src) const { return (int)w->xchg(lane, v, src); }
  int shfl_up(int v, int d) const { return (int)w->xchg(lane, v, lane >= d ? lane - d : lane); }
  int shfl_down(int v, int d) const { return (int)w->xchg(lane, v, lane + d < size ? lane + d : lane); }
  int shfl_xor(int v, int m) const { return (int)w->xchg(lane, v, (lane ^ m) < size ? (lane ^ m) : lane); }
  void sync() const { w->xchg(lane, 0, lane); }
  int max_all(int v) const { const long long* b = w->gather(lane, v); int m = (int)b[0]; for (int i = 1; i < size; ++i) m = m > (int)b[i] ? m : (int)b[i]; return m; }
  int min_all(int v) const { const long long* b = w->gather(lane, v); int m = (int)b[0]; for (int i = 1; i < size; ++i) m = m < (int)b[i] ? m : (int)b[i]; return m; }
};
