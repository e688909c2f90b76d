// warp_emu.h -- TEST INFRASTRUCTURE. Runs the per-environment code of gym_quadruped_b200/csrc/qs_env.cuh on the host:
// a "warp" is 32 fibers (ucontext) run in lock step on the calling thread; __syncwarp / __shfl_sync / __ballot_sync are a
// round-robin hand-over between the fibers and a slot array.  Lets `pytest -m "not gpu"` execute the exact kernel source
// (fp32 and fp64) against the oracle.
//
// Lock step: a lane runs until its next warp barrier and hands over to the next live lane; when control comes back round, every
// other lane has reached the same barrier.  Code between two barriers therefore runs lane after lane, which is one of the
// interleavings the hardware allows for race-free warp code (the kernel source has no other kind: compute-sanitizer racecheck).
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>

#define QS_DEV inline
#define QS_NOINLINE
namespace qs {
struct WarpCtx {
  static constexpr size_t kStack = 512 * 1024;
  ucontext_t sched{}, lane[32]{};
  std::unique_ptr<char[]> stack[32];
  bool done[32] = {};
  uint64_t slots[32] = {};
  std::function<void(int)> body;
  inline void run(std::function<void(int)> fn);
};
inline thread_local WarpCtx* g_ctx = nullptr;
inline thread_local int g_lane = 0;

// hand over to the next live lane (or back to the caller of run() when none is left)
inline void warp_yield(bool finished) {
  WarpCtx* c = g_ctx;
  const int me = g_lane;
  if (finished) c->done[me] = true;
  int nx = me;
  do nx = (nx + 1) & 31; while (c->done[nx] && nx != me);
  if (c->done[nx]) { swapcontext(&c->lane[me], &c->sched); return; }  // every lane has finished
  if (nx == me) return;                                               // the only live lane: the barrier is trivially complete
  g_lane = nx;
  swapcontext(&c->lane[me], &c->lane[nx]);
}
inline void warp_trampoline() {
  WarpCtx* c = g_ctx;
  c->body(g_lane);
  warp_yield(true);
}
inline void WarpCtx::run(std::function<void(int)> fn) {
  body = std::move(fn);
  g_ctx = this;
  for (int i = 0; i < 32; i++) {
    stack[i] = std::make_unique<char[]>(kStack);
    getcontext(&lane[i]);
    lane[i].uc_stack.ss_sp = stack[i].get();
    lane[i].uc_stack.ss_size = kStack;
    lane[i].uc_link = nullptr;
    makecontext(&lane[i], reinterpret_cast<void (*)()>(warp_trampoline), 0);
    done[i] = false;
  }
  g_lane = 0;
  swapcontext(&sched, &lane[0]);
}

inline void syncwarp() { warp_yield(false); }
template <typename T> inline T shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_ctx->slots[g_lane] = raw;
  warp_yield(false);
  uint64_t got = g_ctx->slots[src & 31];
  warp_yield(false);
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
template <typename T> inline T shfl_xor(T v, int o) { return shfl(v, g_lane ^ o); }
inline unsigned ballot(bool p) {
  g_ctx->slots[g_lane] = p ? 1 : 0;
  warp_yield(false);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= unsigned(g_ctx->slots[i] & 1) << i;
  warp_yield(false);
  return m;
}
inline int popc(unsigned x) { return __builtin_popcount(x); }
inline int ctz(unsigned x) { return __builtin_ctz(x); }
inline uint32_t umulhi(uint32_t a, uint32_t b) { return uint32_t((uint64_t(a) * uint64_t(b)) >> 32); }
}  // namespace qs
