// qs_kernel.cuh -- the fused sm_100a kernel family behind libqstep (include/qstep.h).
//
//   env_kernel<real, NCON, MAXDIM, MODE, FEAT>: one environment per warp, WARPS warps per CTA.
//     MODE_STEP    ctrl -> forward dynamics -> Euler -> ALL_OBS / termination      (quadruped_env.py:251-307)
//     MODE_RESET   masked reset: keyframe + noise, lift loop, one step, command / friction resampling (:309-406)
//     MODE_FORWARD forward pass only, dumping accessor tables (mj_forward / mj_fullM / mj_jac users, :543-929)
//     FEAT         compile-time feature switches (qs_env.cuh): specialised step kernels for the BASELINE configurations
//   The robot/scene constants (DModel, ~11 KB) are staged global->shared once per CTA by a single TMA bulk copy
//   (cp.async.bulk + mbarrier) that overlaps with the per-warp state loads; hull vertices stay in global/L2.
//
// Kernels are instantiated in the qs_inst_*.cu translation units (compiled in parallel) and reach the host code of qstep.cu
// through the VariantInfo table declared at the bottom.
#pragma once
// QS_HOST_EMU: the same kernel body compiled by g++ for the host warp emulator (tests/emu/kernel_emu.h supplies threadIdx / blockIdx,
// the warp primitives and plain-memory stand-ins for TMA, mbarrier and the acquire / release accesses) -- test infrastructure only.
#ifndef QS_HOST_EMU
#include <cuda_runtime.h>
#endif

#include <cstddef>
#include <cstdint>

#include "../../include/qstep.h"
#include "qs_env.cuh"

// branch-layout hints: the reset pass, the lift loop, the schedules' resampling and the peer gather are cold in the step kernel
#define QS_LIKELY(x) __builtin_expect(!!(x), 1)
#define QS_UNLIKELY(x) __builtin_expect(!!(x), 0)

namespace qs {

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_FORWARD = 2 };
constexpr int QS_SLOT_ENV_BITS = 20, QS_SLOT_GEN_MASK = 0x7ff;  // queue slot = (generation << 20) | env id; see KParams
constexpr int QS_QUEUE_DEPTH = 8;  // allocated ring entries; QSTEP_RING_DEPTH (2..8, read by qs_create) uses fewer (tests)
constexpr int NCON_MAX = 16;
constexpr int AUX_STRIDE = 324 + 18 + 18 + 216 + 12 + 3 + QS_CONTACT_STRIDE * NCON_MAX + 18 + 18 + 39 + 6 + 3 * 216;  // 1640
constexpr int AUX_OFF_M = 0, AUX_OFF_BIAS = 324, AUX_OFF_PASSIVE = 342, AUX_OFF_JACP = 360, AUX_OFF_FEETPOS = 576, AUX_OFF_COM = 588,
              AUX_OFF_CONTACTS = 591, AUX_OFF_SMOOTH = 591 + QS_CONTACT_STRIDE * NCON_MAX, AUX_OFF_CONSTRAINT = AUX_OFF_SMOOTH + 18,
              AUX_OFF_XPOS = AUX_OFF_CONSTRAINT + 18, AUX_OFF_IMU = AUX_OFF_XPOS + 39, AUX_OFF_JACR = AUX_OFF_IMU + 6,
              AUX_OFF_JACP_DOT = AUX_OFF_JACR + 216, AUX_OFF_JACR_DOT = AUX_OFF_JACP_DOT + 216;

struct KParams {
  const void* dm;      // DModel<real>
  const void* vert;    // Vert4<real>[nvert]
  const void* hf;      // real[nrow*ncol] height-field samples
  const void* boxes;   // DBox<real>[nbox]
  int hm_rows, hm_cols;
  float hm_dx, hm_dy;
  float* hm_out;       // stand-alone ray cast destination [N, rows, cols, 3]
  int num_envs, obs_dim, use_imu, max_iter, env_id_offset, auto_reset;
  int obs_stride;      // floats between consecutive observation rows (obs_dim unless the caller pads its rows, qs_step_host_strided)
  float tol;
  unsigned seed_lo, seed_hi;
  float imu_an, imu_gn, imu_abr, imu_gbr;
  QsBuffers b;
  unsigned* episode;  // per-env reset counter (keys the reset RNG)
  unsigned* tick;     // per-env step counter  (keys the IMU noise RNG)
  // Finish-order queues (MODE_STEP).  A step launch takes its envs from `q_in` (slot -> env id, filled by the previous step launch in
  // the order in which its envs finished; negative = not published yet, see the slot encoding below) and publishes every env it completes to `q_out`.  Because a slot is
  // only handed out once the env behind it has finished its previous step, consecutive step launches may overlap on the device
  // (programmatic dependent launch, QsConfig.pipeline) without any grid-wide dependency: the tail of step t, set by its slowest env,
  // runs next to the head of step t+1.  Envs that finish together are also the ones of similar cost, so a CTA that takes 28
  // consecutive slots gets a homogeneous group (pipelined mode); the serialized mode deals the slots round-robin over the CTAs instead.
  // The queues form a ring of QS_QUEUE_DEPTH entries (launch s reads entry s % DEPTH and fills (s + 1) % DEPTH), so fast envs can run
  // up to DEPTH - 1 launches ahead of a straggler (e.g. an env whose reset lifts it 100 times out of a box) before they wait for it.
  int* q_in; int* q_out;
  unsigned* q_tail;        // monotonic publish counter of q_out
  unsigned q_tail_base;    // its value before this launch's first publish
  int q_contiguous;        // 1: CTA c takes slots [c*W, c*W + W); 0: slot = warp * gridDim + c
  int q_sync;              // 1: launches of this handle may overlap -> acquire / release on the queue slots; 0: the grid boundary orders everything
  // Slot encoding.  A ring entry is read by the launches s, s + depth, s + 2 depth, ... and several of them can be resident at once
  // (all but the oldest spinning on their slots), so a slot value carries the GENERATION of the read it is meant for:
  //   filled:  (gen << 20) | env  (>= 0)          empty: -1 - gen of the read that consumed it  (< 0)
  // gen = (number of earlier reads of this entry) mod 2048.  A consumer only takes a value of its own generation, a producer only
  // overwrites the empty marker of the generation before its own.
  int q_gen_in, q_gen_out;
  // Peer-to-peer observation gather fused into the step (north_star's one collective): with gather_world > 1
  // every warp also stores its finished observation row into the [world * N, D] tensor of EVERY rank (peer-mapped memory, NVLink),
  // and the warp that completes the launch raises this rank's flag on every peer.  `obs` then points into this rank's own tensor.
  float* gather_peers[8];      // base of the parity-selected gathered tensor on rank q (own rank: local memory)
  unsigned* gather_flags[8];   // [world] flags on rank q for this parity; rank r writes entry r
  int gather_world, gather_rank;
  unsigned gather_seq;         // value written to the flags: number of gather steps of this parity so far
  // in-episode schedules (quadruped_env.py:293-305): command resampling ('+reset' types) and external-wrench resampling
  int sch_command_mode, sch_ext_enabled;
  float sch_lin[2], sch_ang[2], sch_ext_lo[6], sch_ext_hi[6];
  unsigned* cmd_epoch; unsigned* ext_epoch;  // per-env draw counters of the two schedules
  const float* ctrl;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  // reset
  const uint8_t* mask;
  const float* in_qpos;
  const float* in_qvel;
  QsResetOptions ro;
  // forward
  float* aux;
#ifdef QS_PROF
  unsigned* prof;  // diagnostic builds only: [N][32] per-env cycle marks (0-15), solver sub-phase cycles (16-23), counters (24-31)
#endif
};
#ifdef QS_PROF
#define QS_MARK(k) do { if (p.prof && lane == 0 && pass == 0) p.prof[size_t(env) * 32 + (k)] = unsigned(clock64() - t_entry); } while (0)
#else
#define QS_MARK(k) do { } while (0)
#endif

#ifndef QS_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// one TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS)
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* mbar) {
  const uint32_t bar = smem_u32(mbar), dst = smem_u32(dst_smem);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src_gmem), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t bar = smem_u32(mbar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
#else
inline void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t*) { std::memcpy(dst_smem, src_gmem, bytes); }
inline void mbar_init(uint64_t*, uint32_t) {}
inline void mbar_wait(uint64_t*, uint32_t) {}
#endif

template <typename real> __device__ __forceinline__ void euler_to_quat(real roll, real pitch, real yaw, real* q) {
  real sr, cr, sp, cp, sy, cy;
  Num<real>::sincos(roll * real(0.5), &sr, &cr);
  Num<real>::sincos(pitch * real(0.5), &sp, &cp);
  Num<real>::sincos(yaw * real(0.5), &sy, &cy);
  q[0] = cr * cp * cy + sr * sp * sy; q[1] = sr * cp * cy - cr * sp * sy; q[2] = cr * sp * cy + sr * cp * sy; q[3] = cr * cp * sy - sr * sp * cy;
}

// Threads per CTA: 28 warps of fp32 envs share one SM (single wave for 4096 envs on 148 SMs, 72 registers/thread);
// the fp64 parity build of the same kernel runs 8 warps per CTA.
template <typename real> struct LaunchCfg { static constexpr int kMaxWarps = 28; };
template <> struct LaunchCfg<double> { static constexpr int kMaxWarps = 8; };

// acquire / release accesses to the finish-order queues (system scope is not needed: producer and consumer are on one GPU)
#ifndef QS_HOST_EMU
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int lane_id() { int v; asm volatile("mov.u32 %0, %%laneid;" : "=r"(v)); return v; }
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define QS_SMEM_DECL extern __shared__ __align__(128) unsigned char smem[]
#else
inline int ld_acquire(const int* p) { return *p; }
inline void st_release(int* p, int v) { *p = v; }
inline unsigned ld_relaxed(const unsigned* p) { return *p; }
inline void st_release_sys(unsigned* p, unsigned v) { *p = v; }
inline int lane_id() { return g_lane; }
inline void launch_dependents() {}
#define QS_SMEM_DECL unsigned char* smem = g_smem
#endif

template <typename real, int NCON, int MAXDIM, int MODE, int FEAT>
__global__ void __launch_bounds__(LaunchCfg<real>::kMaxWarps * 32) env_kernel(const KParams p) {
  using W = WS<real, NCON, MAXDIM>;
  using EnvT = Env<real, NCON, MAXDIM, FEAT>;
  using DM = DModel<real>;
  QS_SMEM_DECL;
  constexpr size_t DM_BYTES = (sizeof(DM) + 127) & ~size_t(127);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + DM_BYTES);
  W* wsbase = reinterpret_cast<W*>(smem + DM_BYTES + 128);
  // canonical warp index broadcast from lane 0: lets the compiler prove it warp-uniform, so the per-warp workspace base lives
  // in a uniform register instead of being re-derived from threadIdx before every shared-memory access
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
  // the lane id is read once through an opaque asm: left to itself the compiler re-materialises `threadIdx.x & 31` with an S2R (a
  // ~25-cycle special-register read) at ~65 places per env-step to save one register
  const int lane = lane_id();
  // The model sits at offset 0 of the dynamic shared memory.  Its base is tied to the (shuffle-produced, hence opaque) warp index so
  // that it lives in a register like the workspace base: otherwise every indexed access to a model table re-derives the shared
  // window base from the CgaCtaId special register (~50 S2R per env-step on address-critical paths).  warp < 32, so the term is 0.
  DM* dm = reinterpret_cast<DM*>(smem + ((warp >> 10) << 4));
  int env = blockIdx.x * nwarp + warp;
#ifdef QS_PROF
  const long long t_entry = clock64();
#endif
  if (threadIdx.x == 0) mbar_init(mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_bulk_load(dm, p.dm, static_cast<uint32_t>(sizeof(DM)), mbar);
  if (MODE == MODE_STEP) {
    // The next step launch may be scheduled as soon as every CTA of this one is resident (or done): whatever env one of its warps
    // waits for is then held by a running warp, so the waits below always end.  Without the pipeline attribute on the next launch
    // this is a no-op.
    launch_dependents();
    // take this warp's env from the finish-order queue filled by the previous step launch (results do not depend on the
    // placement)
    const int slot = p.q_contiguous ? blockIdx.x * nwarp + warp : warp * int(gridDim.x) + int(blockIdx.x);
    env = p.num_envs;
    if (slot < p.num_envs) {
      int e_ = 0;
      if (lane == 0) {
        if (p.q_sync) {
          // only a value of this launch's generation: a later launch that reads the same ring entry may already be spinning here too
          while ((e_ = ld_acquire(p.q_in + slot)) < 0 || (e_ >> QS_SLOT_ENV_BITS) != p.q_gen_in) __nanosleep(200);
          st_release(p.q_in + slot, -1 - p.q_gen_in);  // consumed: free for the producer of the next generation
        } else {
          e_ = p.q_in[slot];  // plain stream order: the previous launch has completed, every slot is filled
          p.q_in[slot] = -1 - p.q_gen_in;
        }
        e_ &= (1 << QS_SLOT_ENV_BITS) - 1;
      }
      env = __shfl_sync(0xffffffffu, e_, 0);
    }
    // The warps of a CTA start their envs together: a convoy that runs through the same code shares its instruction-cache fills
    // (measured: letting every warp start as soon as its own slot is filled costs 45 % throughput on mini_cheetah / flat).
    __syncthreads();
  }

  bool active = env < p.num_envs;
  if (MODE == MODE_RESET && active && p.mask) active = p.mask[env] != 0;
  W& w = wsbase[warp];
  const QsBuffers& B = p.b;

  // ---- state load (overlaps with the TMA copy of the model)
  bool given_state = false;
  if (active) {
    const float* qp = B.qpos + size_t(env) * NQ;
    const float* qv = B.qvel + size_t(env) * NV;
    if (MODE == MODE_RESET && p.in_qpos && p.in_qvel) { qp = p.in_qpos + size_t(env) * NQ; qv = p.in_qvel + size_t(env) * NV; given_state = true; }
    if (lane < NQ) w.qpos[lane] = real(qp[lane]);
    if (lane < NV) { w.qvel[lane] = real(qv[lane]); w.warm[lane] = real(B.qacc_warmstart[size_t(env) * NV + lane]); }
    if (lane < NU) w.ctrl[lane] = (MODE == MODE_STEP) ? real(p.ctrl[size_t(env) * NU + lane]) : real(0);
    if (lane < 6) w.applied[lane] = (MODE == MODE_RESET) ? real(0) : real(B.qfrc_applied[size_t(env) * 6 + lane]);
    if (lane == 0) { w.mu_floor = real(B.friction[2 * env]); w.mu_feet = real(B.friction[2 * env + 1]); }
  }
  mbar_wait(mbar, 0);
  if (!active) return;
  const DM& m = *dm;
  const int e_ttype = (FEAT & FEAT_FLAT) ? 0 : m.terrain_type;
  double* base64 = B.base_pos64 + size_t(env) * 3;
  // "flat here": the scene is the floor plane alone, or the base is out of reach of every other terrain surface (the reference
  // spawns robots over +-10 km in the stairs / ramp scenes).  Then the internal frame is re-centred on the base and the terrain
  // colliders are skipped; near the terrain the frame is the world frame.
  auto flat_at = [&](double x, double y) {
    return e_ttype == 0 || x > double(m.terr_bounds[0]) || x < double(m.terr_bounds[1]) || y > double(m.terr_bounds[2]) || y < double(m.terr_bounds[3]);
  };
  bool flat;
  syncwarp();  // the state rows written above are read across lanes
  {
    const double x0 = (MODE == MODE_RESET && given_state) ? double(w.qpos[0]) : base64[0];
    const double y0 = (MODE == MODE_RESET && given_state) ? double(w.qpos[1]) : base64[1];
    flat = flat_at(x0, y0);
  }
  syncwarp();
  if (lane < 2) {
    // fp64 master copy of the base position (resets scatter envs over +-1e4 m)
    double x = (MODE == MODE_RESET && given_state) ? double(w.qpos[lane]) : base64[lane];
    if (MODE == MODE_RESET && given_state) base64[lane] = x;
    const double o = flat ? rint(x) : 0.0;
    w.org[lane] = o;
    w.qpos[lane] = real(x - o);
  } else if (lane == 2) {
    if (MODE == MODE_RESET && given_state) base64[2] = double(w.qpos[2]); else w.qpos[2] = real(base64[2]);
  }
  syncwarp();

  EnvT e(m, w, reinterpret_cast<const Vert4<real>*>(p.vert), lane);
  e.hf = reinterpret_cast<const real*>(p.hf);
  e.boxes = reinterpret_cast<const DBox<real>*>(p.boxes);
  e.terrain_on = !flat;
  const unsigned env_g = unsigned(env + p.env_id_offset);
  real command[4] = {real(B.command[4 * env]), real(B.command[4 * env + 1]), real(B.command[4 * env + 2]), real(B.command[4 * env + 3])};
  float sim_time = B.sim_time[env];
  int step_count = B.step_count[env];

  // IMU.step (sensors/imu.py:110-139): measurement = truth + bias + noise, bias random walk; counter-based normals.
  // obs_row == nullptr advances the bias walk / counter only.  advance == false (reset passes: the reference's reset() runs mj_step
  // without stepping the sensors, quadruped_env.py:397) writes truth + current bias and leaves the walk and its counter untouched.
  auto imu_step = [&](float* obs_row, bool advance) {
    unsigned tk = p.tick[env];
    if (!advance) {
      if (lane < 3 && obs_row) {
        const float* bias = B.imu_bias + size_t(env) * 6;
        float* io = obs_row + NOBS_BASE;
        io[lane] = float(w.sens[lane]) + bias[lane]; io[3 + lane] = 0.f; io[6 + lane] = bias[lane];
        io[9 + lane] = float(w.sens[3 + lane]) + bias[3 + lane]; io[12 + lane] = 0.f; io[15 + lane] = bias[3 + lane];
      }
      syncwarp();
      return;
    }
    if (lane < 3) {
      uint32_t r[4];
      philox4x32(env_g, tk, unsigned(lane), 0x1A2Bu, p.seed_lo ^ 0x9E3779B9u, p.seed_hi, r);
      const float u1 = fmaxf(u32_to_unit(r[0]), 5.9604645e-8f), u2 = u32_to_unit(r[1]), u3 = fmaxf(u32_to_unit(r[2]), 5.9604645e-8f), u4 = u32_to_unit(r[3]);
      const float ra = sqrtf(-2.f * logf(u1)), rb = sqrtf(-2.f * logf(u3));
      float s1, c1, s2, c2;
      sincosf(6.28318530717958647692f * u2, &s1, &c1);
      sincosf(6.28318530717958647692f * u4, &s2, &c2);
      const float n_acc = ra * c1 * p.imu_an, n_ab = ra * s1 * p.imu_abr, n_gyr = rb * c2 * p.imu_gn, n_gb = rb * s2 * p.imu_gbr;
      float* bias = B.imu_bias + size_t(env) * 6;
      const float ab = bias[lane] + n_ab, gb = bias[3 + lane] + n_gb;
      bias[lane] = ab; bias[3 + lane] = gb;
      if (obs_row) {
        float* io = obs_row + NOBS_BASE;
        io[lane] = float(w.sens[lane]) + ab + n_acc; io[3 + lane] = n_acc; io[6 + lane] = ab;
        io[9 + lane] = float(w.sens[3 + lane]) + gb + n_gyr; io[12 + lane] = n_gyr; io[15 + lane] = gb;
      }
    }
    syncwarp();
    if (lane == 0) p.tick[env] = tk + 1;
  };

  // In-episode schedules, run at the end of every step (never inside a reset): quadruped_env.py:293-305.
  //   '+reset' command types: count the step; when the count reaches the limit drawn with the current command, draw a new
  //   command and a new limit = randint(1000, 3000) (:1046-1072).  External disturbances of type 'reset': same cadence for the
  //   wrench (:1074-1139); the current wrench is then written to qfrc_applied and acts from the next step on (:305).
  auto schedule_update = [&]() {
    if (QS_UNLIKELY((p.sch_command_mode & 8) && lane == 0)) {
      int cnt = B.cmd_count[env] + 1;
      if (QS_UNLIKELY(cnt >= B.cmd_limit[env])) {
        uint32_t r[4];
        const unsigned ce = p.cmd_epoch[env];
        philox4x32(env_g, ce, 0u, 0xC3D0u, p.seed_lo ^ 0x51ED270Bu, p.seed_hi, r);
        p.cmd_epoch[env] = ce + 1;
        const real vn = real(double(p.sch_lin[0]) + (double(p.sch_lin[1]) - double(p.sch_lin[0])) * double(u32_to_unit(r[0])));
        real hx = 1, hy = 0, vnorm = vn;
        if (p.sch_command_mode & 2) { real ang = real(-3.14159265358979323846 + 2 * 3.14159265358979323846 * double(u32_to_unit(r[1]))); Num<real>::sincos(ang, &hy, &hx); }
        if (!(p.sch_command_mode & 3)) vnorm = 0;  // 'human': zero speed (:1058-1061)
        const real yr = (p.sch_command_mode & 4) ? real(double(p.sch_ang[0]) + (double(p.sch_ang[1]) - double(p.sch_ang[0])) * double(u32_to_unit(r[2]))) : real(0);
        B.command[4 * env] = float(vnorm * hx); B.command[4 * env + 1] = float(vnorm * hy); B.command[4 * env + 2] = 0.f; B.command[4 * env + 3] = float(yr);
        B.cmd_limit[env] = 1000 + int(u32_to_unit(r[3]) * 2000.f);
        cnt = 0;
      }
      B.cmd_count[env] = cnt;
    }
    if (QS_UNLIKELY(p.sch_ext_enabled)) {
      int due = 0;
      if (lane == 0) {
        const int cnt = B.ext_count[env] + 1;
        due = cnt >= B.ext_limit[env];
        B.ext_count[env] = due ? 0 : cnt;
      }
      due = __shfl_sync(0xffffffffu, due, 0);
      if (QS_UNLIKELY(due)) {
        const unsigned ee = p.ext_epoch[env];
        if (lane < 7) {
          uint32_t r[4];
          philox4x32(env_g, ee, unsigned(lane), 0xD157u, p.seed_lo ^ 0x7F4A7C15u, p.seed_hi, r);
          const float u = u32_to_unit(r[0]);
          if (lane < 6) B.ext_wrench[size_t(env) * 6 + lane] = float(double(p.sch_ext_lo[lane]) + (double(p.sch_ext_hi[lane]) - double(p.sch_ext_lo[lane])) * double(u));
          else B.ext_limit[env] = 1000 + int(u * 2000.f);
        }
        syncwarp();
        if (lane == 0) p.ext_epoch[env] = ee + 1;
      }
      syncwarp();
      if (lane < 6) B.qfrc_applied[size_t(env) * 6 + lane] = B.ext_wrench[size_t(env) * 6 + lane];
    }
  };

  // One pass = one "mj_step" with its env-side bookkeeping.  A reset is the same pass preceded by state sampling and the
  // lift loop; MODE_STEP with auto_reset runs a second (reset) pass for envs that just terminated, in the same warp.
  bool resetting = (MODE == MODE_RESET);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    unsigned status = 0;
    int lift_phase = 2;
    real u_late[5] = {0, 0, 0, 0, 0};
    if (MODE != MODE_FORWARD && (MODE == MODE_RESET ? resetting : QS_UNLIKELY(resetting))) {
      const QsResetOptions& ro = p.ro;
      const unsigned ep = p.episode[env];
      real* u = w.obs;  // scratch for the uniforms (this storage is recycled by the solver later on)
      if (lane < 10) {
        uint32_t r[4];
        philox4x32(env_g, ep, unsigned(lane), 0x5EEDu, p.seed_lo, p.seed_hi, r);
        for (int i = 0; i < 4; i++) u[4 * lane + i] = real(u32_to_unit(r[i]));
        if (lane == 9) {  // two 53-bit uniforms for the fp64 base xy
          w.tmpd[0] = (double(r[0]) * 4294967296.0 + double(r[1])) * (1.0 / 18446744073709551616.0);
          w.tmpd[1] = (double(r[2]) * 4294967296.0 + double(r[3])) * (1.0 / 18446744073709551616.0);
        }
      }
      syncwarp();
      for (int i = 0; i < 5; i++) u_late[i] = u[26 + i];  // consumed after the step
      if (!given_state) {
        const real dq = (lane < NJ && ro.randomize) ? real(-ro.angle_sweep + 2 * ro.angle_sweep * double(u[lane])) : real(0);
        const real dv = (lane < NJ && ro.randomize) ? real(-ro.vel_sweep + 2 * ro.vel_sweep * double(u[12 + lane])) : real(0);
        const real roll = real(-ro.roll_sweep + 2 * ro.roll_sweep * double(u[24])), pitch = real(-ro.pitch_sweep + 2 * ro.pitch_sweep * double(u[25]));
        syncwarp();
        if (lane < NQ) w.qpos[lane] = m.key_qpos[lane];
        if (lane < NV) w.qvel[lane] = 0;
        syncwarp();
        double bx = double(m.key_qpos[0]), by = double(m.key_qpos[1]);
        if (ro.randomize) {
          if (lane < NJ) { w.qpos[7 + lane] += dq; w.qvel[6 + lane] += dv; }
          // np.random.uniform(limits[0], limits[1]) = lo + (hi - lo) * u with lo = x_max, hi = x_min (quadruped_env.py:352-356)
          bx = double(m.terrain_limits[0]) + (double(m.terrain_limits[1]) - double(m.terrain_limits[0])) * w.tmpd[0];
          by = double(m.terrain_limits[2]) + (double(m.terrain_limits[3]) - double(m.terrain_limits[2])) * w.tmpd[1];
          if (lane == 0) {
            const real yaw = real(atan2(-by, -bx));  // angle_between_vectors(xy, 0) math_utils.py:50-51
            real q[4];
            euler_to_quat(roll, pitch, yaw, q);
            for (int i = 0; i < 4; i++) w.qpos[3 + i] = q[i];
            w.qpos[2] = real(ro.hip_height);
          }
        }
        syncwarp();
        flat = flat_at(bx, by);
        e.terrain_on = !flat;
        if (lane == 0) {
          const double ox = flat ? rint(bx) : 0.0, oy = flat ? rint(by) : 0.0;
          w.org[0] = ox; w.org[1] = oy;
          w.qpos[0] = real(bx - ox); w.qpos[1] = real(by - oy);
        }
        syncwarp();
        lift_phase = flat ? 0 : 1;  // the lift loop shares the position stage below (one copy of the collider code per kernel)
      }
      // zero ctrl / applied wrench / warm start / clock (quadruped_env.py:332-335, :394-395)
      if (lane < NV) w.warm[lane] = 0;
      if (lane < NU) w.ctrl[lane] = 0;
      if (lane < 6) w.applied[lane] = 0;
      sim_time = 0.f;
      syncwarp();
    }

    // ---- position stage; a random reset first lifts the robot until no foot (calf-body) contact is left, quadruped_env.py:376-388
    //   phase 0  flat floor, first pass: raising the base shifts every floor distance by exactly the lift, so after one full
    //            collision pass the loop reduces to a scalar recurrence on the calf-body contact distances (same iterates as
    //            re-running the collision stage) -- unless a calf box is in contact (a box keeps at most 4 corners: not closed
    //            under lifting) or the contact buffer overflowed; then phase 1 takes over from the start
    //   phase 1  kinematics + collision restricted to the calf geoms, raise by 1.1 x the deepest penetration, at most 100 times
    //   phase 2  the forward pass proper
    QS_MARK(1);
    {
      bool cleared = lift_phase == 2;
      int c = 0;
#pragma unroll 1
      for (;;) {
        e.calf_only = lift_phase == 1;
        e.kinematics();
        if (lift_phase == 2) { e.com_inertia(); e.cdof(); }
        e.collide_floor();
        if (QS_LIKELY(lift_phase == 2)) break;
        if (lift_phase == 0) {
          real d = Num<real>::big, mg = 0;
          bool calf = false, boxy = false;
          if (lane < w.ncon) {
            const int info = w.c_info[lane], bdy = (info >> 8) & 0xff, g = info & 0xff;
            calf = bdy >= 2 && (bdy - 2) % 3 == 2;
            d = w.c_dist[lane]; mg = m.geom_margin[g];
            boxy = calf && m.geom_type[g] == GEOM_BOX;
          }
          lift_phase = 1;
          if (!w.overflow && qs::ballot(boxy) == 0) {
            real lift = 0;
#pragma unroll 1
            for (int k = 0; k <= 100; k++) {
              const bool in = calf && !(d + lift > mg);
              const real pen = warp_max(in ? Num<real>::abs(d + lift) : real(0));
              if (qs::ballot(in) == 0) { cleared = true; break; }
              if (k == 100) break;
              lift += pen * real(1.1);
            }
            if (lane == 0) w.qpos[2] += lift;
            syncwarp();
            lift_phase = 2;
          }
          continue;
        }
        // calf-body contacts as DETECTED by the collision pass, not as stored: the 16-slot contact buffer may have dropped the deepest
        const bool any = qs::ballot(e.cm_acc != 0) != 0;
        const real pen = warp_max(e.pen_acc);
        if (!any) { cleared = true; lift_phase = 2; continue; }
        if (c == 100) { lift_phase = 2; continue; }
        if (lane == 0) w.qpos[2] += pen * real(1.1);
        syncwarp();
        c++;
#ifdef QS_PROF
        if (p.prof && lane == 0) { p.prof[size_t(env) * 32 + 29] = unsigned(c); p.prof[size_t(env) * 32 + 30] = __float_as_uint(float(pen)); }
#endif
      }
      if (!cleared) status |= 8u;
    }
    QS_MARK(2);
    typename EnvT::Flags fl = e.flags();  // contact masks depend on the collision stage only
    if (MODE == MODE_STEP && p.auto_reset && !resetting) {
      // Same-step auto-reset returns the post-reset state / observation of an env that terminates, so once the collision stage has
      // found a contact that terminates the episode (quadruped_env.py:1228-1248) the rest of this step cannot reach any output:
      // raise the flags, keep the IMU bias walk in step, and go straight to the reset pass.
      if (QS_UNLIKELY(fl.invalid_mask != 0)) {
        if (lane == 0) {
          if (p.reward) p.reward[env] = 0.f;
          if (p.terminated) p.terminated[env] = 1;
          if (p.truncated) p.truncated[env] = 0;
        }
        if (!(FEAT & FEAT_NO_IMU) && p.use_imu) imu_step(nullptr, true);
        schedule_update();
        QS_MARK(5);
        resetting = true;
        given_state = false;
        syncwarp();
        continue;
      }
    }
    if (MODE == MODE_FORWARD && p.aux) {
      float* a = p.aux + size_t(env) * AUX_STRIDE;  // body poses are only valid until the solver recycles their storage
      for (int it = lane; it < 39; it += 32) a[AUX_OFF_XPOS + it] = float(w.kin.xpos[1 + it / 3][it % 3] + (it % 3 < 2 ? real(w.org[it % 3]) : real(0)));
      syncwarp();
      e.bias_out = a + AUX_OFF_BIAS;
    }
#ifdef QS_PROF
    e.bias_and_smooth(); e.mass_matrix(); e.make_constraints();
    QS_MARK(3);
    e.solve(p.max_iter, real(p.tol));
    if (e.has_imu()) e.sensors();
    QS_MARK(4);
    if (p.prof && lane == 0 && pass == 0) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned* pr = p.prof + size_t(env) * 32;
      pr[24] = unsigned(e.solver_iter); pr[25] = unsigned(e.ls_evals); pr[26] = unsigned(w.ncon); pr[27] = smid; pr[28] = unsigned(warp);
      pr[0] = unsigned(t_entry & 0xffffffffll);
      for (int i = 0; i < 8; i++) pr[16 + i] = e.tacc[i];
    }
#else
    if (MODE == MODE_FORWARD && p.aux) {
      // body velocities / cdof_dot live in storage that the constraint stage recycles: export the Jacobian tables in between
      e.bias_and_smooth();
      float* a = p.aux + size_t(env) * AUX_STRIDE;
      e.dump_jacobian_tables(a + AUX_OFF_JACR, a + AUX_OFF_JACP_DOT, a + AUX_OFF_JACR_DOT);
      syncwarp();
      e.mass_matrix(); e.make_constraints(); e.solve(p.max_iter, real(p.tol));
      if (e.has_imu()) e.sensors();
    } else {
      e.forward_dynamics(p.max_iter, real(p.tol));
    }
#endif

    if (MODE == MODE_FORWARD) {
      if (lane < NV) B.qacc[size_t(env) * NV + lane] = float(w.qacc[lane]);
      if (p.aux) {
        float* a = p.aux + size_t(env) * AUX_STRIDE;
        for (int it = lane; it < 324; it += 32) {
          const int i = it / 18, j = it % 18;
          real v = 0;
          if (i < 6 && j < 6) v = w.Mbb[i][j];
          else if (i >= 6 && j < 6) v = w.Mlb[(i - 6) / 3][(i - 6) % 3][j];
          else if (i < 6 && j >= 6) v = w.Mlb[(j - 6) / 3][(j - 6) % 3][i];
          else if ((i - 6) / 3 == (j - 6) / 3) v = w.Mll[(i - 6) / 3][(i - 6) % 3][(j - 6) % 3];
          a[AUX_OFF_M + it] = float(v);
        }
        if (lane < NV) {
          a[AUX_OFF_PASSIVE + lane] = float(-m.dof_damping[lane] * w.qvel[lane]);
          a[AUX_OFF_SMOOTH + lane] = float(w.fsm[lane]);
          a[AUX_OFF_CONSTRAINT + lane] = float(w.fcon[lane]);
        }
        for (int it = lane; it < 216; it += 32) {
          const int l = it / 54, i = (it % 54) / 18, d = it % 18;
          real v = 0;
          if (d < 6 || (d - 6) / 3 == l) {
            const real off[3] = {w.footpos[l][0] - w.com[0], w.footpos[l][1] - w.com[1], w.footpos[l][2] - w.com[2]};
            real cr[3];
            cross3(cr, w.cdof[d], off);
            v = w.cdof[d][3 + i] + cr[i];
          }
          a[AUX_OFF_JACP + it] = float(v);
        }
        if (lane < 12) a[AUX_OFF_FEETPOS + lane] = float(w.footpos[lane / 3][lane % 3] + (lane % 3 < 2 ? real(w.org[lane % 3]) : real(0)));
        if (lane < 3) a[AUX_OFF_COM + lane] = float(w.com[lane] + (lane < 2 ? real(w.org[lane]) : real(0)));
        if (lane < 6) a[AUX_OFF_IMU + lane] = e.has_imu() ? float(w.sens[lane]) : 0.f;
        for (int c = lane; c < NCON; c += 32) {
          float* o = a + AUX_OFF_CONTACTS + QS_CONTACT_STRIDE * c;
          if (c < w.ncon) {
            const int info = w.c_info[c], dim = (info >> 16) & 0xff;
            o[0] = float(w.c_dist[c]);
            o[1] = float(w.c_pos[c][0] + real(w.org[0])); o[2] = float(w.c_pos[c][1] + real(w.org[1])); o[3] = float(w.c_pos[c][2]);
            real t2[3];
            cross3(t2, w.c_frame[c], w.c_frame[c] + 3);
            for (int i = 0; i < 6; i++) o[4 + i] = float(w.c_frame[c][i]);
            for (int i = 0; i < 3; i++) o[10 + i] = float(t2[i]);
            for (int i = 0; i < 3; i++) o[13 + i] = (i < dim) ? float(w.c_F[c][i]) : 0.f;
            o[16] = float(info & 0xff); o[17] = float((info >> 8) & 0xff); o[18] = float(w.c_fri[c][0]); o[19] = float(dim);
          } else {
            for (int i = 0; i < QS_CONTACT_STRIDE; i++) o[i] = 0.f;
          }
        }
      }
      if (lane == 0) {
        B.ncon[env] = w.ncon;
        B.solver_iter[env] = e.solver_iter | (e.ls_evals << 8);
        B.invalid_body_mask[2 * env] = uint8_t(fl.invalid_mask & 0xff); B.invalid_body_mask[2 * env + 1] = uint8_t((fl.invalid_mask >> 8) & 0xff);
      }
      return;
    }

    // ---- integrate, then env-side bookkeeping
    e.integrate(base64);
    QS_MARK(6);
    fl.out_of_bounds = e.out_of_bounds();  // bounds are tested on the post-step base position (:1252-1256)
    // A non-finite state also ends the episode (status bit0): the engine warns and resets its data in that case; here the reset
    // is the caller's (or the auto-reset pass's), so that one bad env cannot stay bad for the rest of a rollout.
    bool finite_state = true;
    if (lane < NQ) finite_state = isfinite(w.qpos[lane]);
    if (lane < NV) finite_state = finite_state && isfinite(w.qvel[lane]);
    if (QS_UNLIKELY(qs::ballot(!finite_state) != 0)) status |= 1u;
    const bool terminated = fl.invalid_mask != 0 || fl.out_of_bounds || (status & 1u);
    sim_time += float(m.timestep);
    step_count = resetting ? 0 : step_count + 1;
    if (resetting) {
      // command + friction resampling happen after the step inside reset (:397-404)
      const QsResetOptions& ro = p.ro;
      const real vn = real(ro.lin_vel_range[0] + (ro.lin_vel_range[1] - ro.lin_vel_range[0]) * double(u_late[0]));
      real hx = 1, hy = 0, vnorm = vn;
      if (ro.command_mode & 2) { real ang = real(-3.14159265358979323846 + 2 * 3.14159265358979323846 * double(u_late[1])); Num<real>::sincos(ang, &hy, &hx); }
      if (!(ro.command_mode & 3)) vnorm = 0;  // 'human'
      command[0] = vnorm * hx; command[1] = vnorm * hy; command[2] = 0;
      command[3] = (ro.command_mode & 4) ? real(ro.ang_vel_range[0] + (ro.ang_vel_range[1] - ro.ang_vel_range[0]) * double(u_late[2])) : real(0);
      const float mu = float(ro.friction_range[0] + (ro.friction_range[1] - ro.friction_range[0]) * double(u_late[3]));
      if (lane < 4) B.command[4 * env + lane] = float(command[lane]);
      if (lane < 2) B.friction[2 * env + lane] = mu;
      if (lane == 0) {
        p.episode[env] = p.episode[env] + 1; w.mu_floor = real(mu); w.mu_feet = real(mu);
        if (ro.command_mode & 8) { B.cmd_count[env] = 0; B.cmd_limit[env] = 1000 + int(float(u_late[4]) * 2000.f); }  // :1068-1070
      }
      if (lane < 6) B.qfrc_applied[size_t(env) * 6 + lane] = 0.f;
    }
    syncwarp();
    e.pack_obs(command, fl.contact_mask);
    QS_MARK(7);

    // ---- write back
    if (w.overflow) status |= 2u;
    if (e.solver_maxed) status |= 4u;
    if (lane < NQ) B.qpos[size_t(env) * NQ + lane] = (lane < 3) ? float(base64[lane]) : float(w.qpos[lane]);
    if (lane < NV) {
      B.qvel[size_t(env) * NV + lane] = float(w.qvel[lane]);
      B.qacc[size_t(env) * NV + lane] = float(w.qacc[lane]);
      B.qacc_warmstart[size_t(env) * NV + lane] = float(w.qacc[lane]);
    }
    float* obs = p.obs ? p.obs + size_t(env) * p.obs_stride : nullptr;
    if (obs) {
      // The staged row leaves as a scalar head up to 16-B alignment, a float4 body and a scalar tail: rows bound for mapped host
      // memory cross PCIe in 16-B stores (measured 45 GB/s against 38 GB/s for 4-B stores, scripts/micro/zc_write.cu).
      const int head = (4 - int((reinterpret_cast<size_t>(obs) >> 2) & 3)) & 3;
      if (lane < head) obs[lane] = float(w.obs[lane]);
      const int nvec = (NOBS_BASE - head) >> 2;
      for (int v = lane; v < nvec; v += 32) {
        const int i = head + 4 * v;
        *reinterpret_cast<float4*>(obs + i) = make_float4(float(w.obs[i]), float(w.obs[i + 1]), float(w.obs[i + 2]), float(w.obs[i + 3]));
      }
      const int done = head + 4 * nvec;
      if (lane < NOBS_BASE - done) obs[done + lane] = float(w.obs[done + lane]);
    }
    if (!(FEAT & FEAT_NO_HM) && obs && p.hm_rows > 0) {
      // sensors/heightmap columns: grid around the post-step base position / heading (heightmap.py:106-169)
      real qq[4] = {w.qpos[3], w.qpos[4], w.qpos[5], w.qpos[6]}, Rn[9];
      quat_normalize(qq);
      quat_to_mat(Rn, qq);
      const real ctr[3] = {w.qpos[0], w.qpos[1], w.qpos[2]};
      e.heightmap(ctr, Num<real>::atan2(Rn[3], Rn[0]), p.hm_rows, p.hm_cols, real(p.hm_dx), real(p.hm_dy), real(w.org[0]), real(w.org[1]),
                  obs + NOBS_BASE + (p.use_imu ? QS_NOBS_IMU : 0));
    }
    if (!(FEAT & FEAT_NO_IMU) && p.use_imu) imu_step(obs, !resetting);
    if (MODE == MODE_STEP && QS_UNLIKELY(p.gather_world > 1) && obs && !(p.auto_reset && !resetting && terminated)) {
      // final observation row of this env (a terminated env under auto-reset sends its post-reset row from the second pass):
      // straight to the gathered tensor of every peer, 16-B stores where the destination allows
      syncwarp();
      const size_t row = (size_t(p.gather_rank) * p.num_envs + env) * p.obs_stride;
      auto val = [&](int i) { return i < NOBS_BASE ? float(w.obs[i]) : __ldcg(obs + i); };
      for (int q = 0; q < p.gather_world; q++) {
        if (q == p.gather_rank) continue;
        float* dst = p.gather_peers[q] + row;
        const int head = (4 - int((reinterpret_cast<size_t>(dst) >> 2) & 3)) & 3;
        if (lane < head) dst[lane] = val(lane);
        const int nvec = (p.obs_dim - head) >> 2;
        for (int v = lane; v < nvec; v += 32) {
          const int i = head + 4 * v;
          *reinterpret_cast<float4*>(dst + i) = make_float4(val(i), val(i + 1), val(i + 2), val(i + 3));
        }
        const int done = head + 4 * nvec;
        if (lane < p.obs_dim - done) dst[done + lane] = val(done + lane);
      }
    }
    if (MODE == MODE_STEP && !resetting) schedule_update();
    if (lane == 0) {
      B.sim_time[env] = sim_time;
      B.step_count[env] = step_count;
      B.status[env] = uint8_t(status);
      B.ncon[env] = w.ncon;
      B.solver_iter[env] = e.solver_iter | (e.ls_evals << 8);  // low byte: Newton iterations, upper bits: line-search evaluations
      B.invalid_body_mask[2 * env] = uint8_t(fl.invalid_mask & 0xff); B.invalid_body_mask[2 * env + 1] = uint8_t((fl.invalid_mask >> 8) & 0xff);
      if (MODE == MODE_STEP && !resetting) {
        if (p.reward) p.reward[env] = 0.f;  // _compute_reward, quadruped_env.py:1141-1144
        if (p.terminated) p.terminated[env] = uint8_t(terminated);
        if (p.truncated) p.truncated[env] = 0;
      }
    }
    QS_MARK(5);
    // in-kernel auto-reset: the warp of an env that just terminated goes round once more as a reset pass
    if (QS_LIKELY(MODE != MODE_STEP || !p.auto_reset || resetting || !terminated)) break;
    resetting = true;
    given_state = false;
    e.ls_evals = 0;
    syncwarp();
  }
  if (MODE == MODE_STEP) {
    // publish: this env may now be stepped again.  The warp's global writes are ordered before the release by the warp barrier
    // (all lanes' stores happen-before lane 0's fence) -- the consumer's acquire load pairs with it.
    __syncwarp();
    if (lane == 0) {
      if (p.gather_world > 1) __threadfence_system();  // this warp's peer stores are performed before it counts as finished
      if (p.q_sync) {
        // The publish counter of a ring entry is shared by every launch that fills it (s, s + depth, ...): wait until the previous
        // one has published all of its envs, i.e. until the counter has reached this launch's base.  Only an env that lags `depth`
        // launches behind (a reset that lifts the robot 100 times) ever makes a fast env wait here.
        unsigned t_;
        do {
          t_ = ld_relaxed(p.q_tail);
          if (int(t_ - p.q_tail_base) < 0) __nanosleep(200);
        } while (int(t_ - p.q_tail_base) < 0);
      }
      const unsigned pos = atomicAdd(p.q_tail, 1u) - p.q_tail_base;
      const int filled = (p.q_gen_out << QS_SLOT_ENV_BITS) | env;
      if (p.q_sync) {
        // wait for the consumer of this slot's previous generation (far behind, or not even served yet by its own producer)
        const int want = -1 - ((p.q_gen_out - 1) & QS_SLOT_GEN_MASK);
        while (ld_acquire(p.q_out + pos) != want) __nanosleep(200);
        st_release(p.q_out + pos, filled);
      } else {
        p.q_out[pos] = filled;
      }
      if (p.gather_world > 1 && pos == unsigned(p.num_envs) - 1u) {
        // last env of the launch: every other warp fenced before its increment, so all rows of this rank are on their way before
        // the flag; release at system scope, the peers' wait kernel acquires it
        __threadfence_system();
        for (int q = 0; q < p.gather_world; q++)
          st_release_sys(p.gather_flags[q] + p.gather_rank, p.gather_seq);
      }
    }
  }
#ifdef QS_PROF
  if (p.prof && lane == 0) p.prof[size_t(env) * 32 + 15] = unsigned(clock64() - t_entry);
#endif
}

#ifndef QS_HOST_EMU
// HeightMap.update_height_map for every env (sensors/heightmap.py:106-169): one warp per env, rays spread over the lanes
template <typename real>
__global__ void __launch_bounds__(256) raycast_kernel(const KParams p) {
  const int lane = threadIdx.x & 31, env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const DModel<real>& m = *reinterpret_cast<const DModel<real>*>(p.dm);
  using W = WS<real, NCON_MAX, 3>;
  Env<real, NCON_MAX, 3, 0> e(m, *reinterpret_cast<W*>(smem), reinterpret_cast<const Vert4<real>*>(p.vert), lane);  // workspace is never touched
  e.hf = reinterpret_cast<const real*>(p.hf);
  e.boxes = reinterpret_cast<const DBox<real>*>(p.boxes);
  const double* b64 = p.b.base_pos64 + size_t(env) * 3;
  const float* qp = p.b.qpos + size_t(env) * NQ;
  const bool flat = m.terrain_type == 0 || b64[0] > double(m.terr_bounds[0]) || b64[0] < double(m.terr_bounds[1]) ||
                    b64[1] > double(m.terr_bounds[2]) || b64[1] < double(m.terr_bounds[3]);
  e.terrain_on = !flat;
  const double ox = flat ? rint(b64[0]) : 0.0, oy = flat ? rint(b64[1]) : 0.0;
  real qq[4] = {real(qp[3]), real(qp[4]), real(qp[5]), real(qp[6])}, R[9];
  quat_normalize(qq);
  quat_to_mat(R, qq);
  const real ctr[3] = {real(b64[0] - ox), real(b64[1] - oy), real(b64[2])};
  e.heightmap(ctr, Num<real>::atan2(R[3], R[0]), p.hm_rows, p.hm_cols, real(p.hm_dx), real(p.hm_dy), real(ox), real(oy),
              p.hm_out + size_t(env) * p.hm_rows * p.hm_cols * 3);
}using KernelFn = void (*)(const KParams);

// one compiled kernel variant: precision (0 fp32, 1 fp64), contact dimension of the workspace, asserted features
struct VariantInfo {
  const char* name;
  int precision, maxdim, feat;
  KernelFn step, reset, forward, raycast;  // specialised variants carry `step` only
  size_t ws_bytes, dm_bytes;
  int max_warps;
};

template <typename real, int MAXDIM, int FEAT, bool ALL_MODES> VariantInfo make_variant(const char* name) {
  VariantInfo v{};
  v.name = name; v.precision = sizeof(real) == 4 ? 0 : 1; v.maxdim = MAXDIM; v.feat = FEAT;
  v.step = env_kernel<real, NCON_MAX, MAXDIM, MODE_STEP, FEAT>;
  if constexpr (ALL_MODES) {
    v.reset = env_kernel<real, NCON_MAX, MAXDIM, MODE_RESET, FEAT>;
    v.forward = env_kernel<real, NCON_MAX, MAXDIM, MODE_FORWARD, FEAT>;
  }
  v.ws_bytes = sizeof(WS<real, NCON_MAX, MAXDIM>);
  v.dm_bytes = (sizeof(DModel<real>) + 127) & ~size_t(127);
  v.max_warps = LaunchCfg<real>::kMaxWarps;
  return v;
}

#endif  // QS_HOST_EMU

}  // namespace qs
