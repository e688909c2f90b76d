// warp_emu.h -- TEST INFRASTRUCTURE. Runs the per-environment code of gym_quadruped_b200/csrc/qs_env.cuh on the host:
// a "warp" is 32 std::threads in lock step; __syncwarp / __shfl_sync / __ballot_sync are emulated with a barrier and a
// slot array.  Lets `pytest -m "not gpu"` execute the exact kernel source (fp32 and fp64) against the oracle.
#pragma once
#include <barrier>
#include <cstdint>
#include <cstring>

#define QS_DEV inline
#define QS_NOINLINE
namespace qs {
struct WarpCtx {
  std::barrier<> bar{32};
  uint64_t slots[32];
};
inline thread_local WarpCtx* g_ctx = nullptr;
inline thread_local int g_lane = 0;
inline void syncwarp() { g_ctx->bar.arrive_and_wait(); }
template <typename T> inline T shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  g_ctx->slots[g_lane] = raw;
  g_ctx->bar.arrive_and_wait();
  uint64_t got = g_ctx->slots[src & 31];
  g_ctx->bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
template <typename T> inline T shfl_xor(T v, int o) { return shfl(v, g_lane ^ o); }
inline unsigned ballot(bool p) {
  g_ctx->slots[g_lane] = p ? 1 : 0;
  g_ctx->bar.arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= unsigned(g_ctx->slots[i] & 1) << i;
  g_ctx->bar.arrive_and_wait();
  return m;
}
inline int popc(unsigned x) { return __builtin_popcount(x); }
inline int ctz(unsigned x) { return __builtin_ctz(x); }
inline uint32_t umulhi(uint32_t a, uint32_t b) { return uint32_t((uint64_t(a) * uint64_t(b)) >> 32); }
}  // namespace qs
