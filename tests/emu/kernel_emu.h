// kernel_emu.h -- TEST INFRASTRUCTURE.  Host stand-ins for the CUDA constructs used by gym_quadruped_b200/csrc/qs_kernel.cuh so that
// the step / reset / forward KERNEL BODY (reset noise, lift loop, auto-reset pass, in-kernel schedules, flags, write-back) runs on
// the warp emulator of warp_emu.h: one CTA = one warp = 32 fibers.  Include before qs_kernel.cuh with QS_HOST_EMU defined.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "warp_emu.h"

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(x)
#define __align__(x)

namespace qs {
struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmuDim3 blockIdx, blockDim, gridDim;
inline thread_local unsigned char* g_smem = nullptr;
// threadIdx.x depends on the running fiber
struct EmuThreadIdx { struct X { operator unsigned() const { return unsigned(g_lane); } } x; };
inline EmuThreadIdx threadIdx;

template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return shfl(v, src); }
inline void __syncthreads() { syncwarp(); }  // one warp per emulated CTA
inline void __syncwarp() { syncwarp(); }
inline void __nanosleep(unsigned) {}
inline void __threadfence_system() {}
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned old = *p; *p = old + v; return old; }
inline float __ldcg(const float* p) { return *p; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
using std::isfinite;
}  // namespace qs
