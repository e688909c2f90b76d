// qstep.cu -- libqstep: the C-ABI of include/qstep.h on top of one fused sm_100a kernel family.
//
//   env_kernel<real, NCON, MAXDIM, MODE>: one environment per warp, WARPS warps per CTA.
//     MODE_STEP    ctrl -> forward dynamics -> Euler -> ALL_OBS / termination      (quadruped_env.py:251-307)
//     MODE_RESET   masked reset: keyframe + noise, lift loop, one step, command / friction resampling (:309-406)
//     MODE_FORWARD forward pass only, dumping accessor tables (mj_forward / mj_fullM / mj_jac users, :543-929)
//   The robot/scene constants (DModel, ~10 KB) are staged global->shared once per CTA by a single TMA bulk copy
//   (cp.async.bulk + mbarrier) that overlaps with the per-warp state loads; hull vertices stay in global/L2.
//
// No torch types cross this boundary; PyTorch only owns the device buffers whose pointers arrive in QsBuffers.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "qs_host_model.h"

namespace {
using namespace qs;

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_FORWARD = 2 };
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
  float tol;
  unsigned seed_lo, seed_hi;
  float imu_an, imu_gn, imu_abr, imu_gbr;
  QsBuffers b;
  unsigned* episode;  // per-env reset counter (keys the reset RNG)
  unsigned* tick;     // per-env step counter  (keys the IMU noise RNG)
  // straggler-aware placement (MODE_STEP): envs that were contact-rich / slow to solve in the previous step are listed first
  const int* sched_in; const int* sched_in_cnt;  // [2N] heavy | light lists and their two counters (nullptr: identity placement)
  int* sched_out; int* sched_out_cnt; int* sched_zero_cnt;
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

template <typename real, int NCON, int MAXDIM, int MODE>
__global__ void __launch_bounds__(LaunchCfg<real>::kMaxWarps * 32) env_kernel(const KParams p) {
  using W = WS<real, NCON, MAXDIM>;
  using DM = DModel<real>;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr size_t DM_BYTES = (sizeof(DM) + 127) & ~size_t(127);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + DM_BYTES);
  W* wsbase = reinterpret_cast<W*>(smem + DM_BYTES + 128);
  // canonical warp index broadcast from lane 0: lets the compiler prove it warp-uniform, so the per-warp workspace base lives
  // in a uniform register instead of being re-derived from threadIdx before every shared-memory access
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), nwarp = blockDim.x >> 5;
  // the lane id is read once through an opaque asm: left to itself the compiler re-materialises `threadIdx.x & 31` with an S2R (a
  // ~25-cycle special-register read) at ~65 places per env-step to save one register
  int lane_reg;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_reg));
  const int lane = lane_reg;
  // The model sits at offset 0 of the dynamic shared memory.  Its base is tied to the (shuffle-produced, hence opaque) warp index so
  // that it lives in a register like the workspace base: otherwise every indexed access to a model table re-derives the shared
  // window base from the CgaCtaId special register (~50 S2R per env-step on address-critical paths).  warp < 32, so the term is 0.
  DM* dm = reinterpret_cast<DM*>(smem + ((warp >> 10) << 4));
  if (MODE == MODE_STEP && p.sched_zero_cnt && blockIdx.x == 0 && threadIdx.x < 2) p.sched_zero_cnt[threadIdx.x] = 0;
  int env = blockIdx.x * nwarp + warp;
  if (MODE == MODE_STEP && p.sched_in) {
    // Heavy envs first, dealt round-robin over the CTAs and onto the highest warp ids (the issue arbiter favours those): every SM
    // gets the same share of likely stragglers and starts them early.  Results do not depend on the placement.
    const int k = (nwarp - 1 - warp) * gridDim.x + blockIdx.x;
    const int n_heavy = p.sched_in_cnt[0];
    env = k < p.num_envs ? (k < n_heavy ? p.sched_in[k] : p.sched_in[p.num_envs + (k - n_heavy)]) : p.num_envs;
  }

#ifdef QS_PROF
  const long long t_entry = clock64();
#endif
  if (threadIdx.x == 0) mbar_init(mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_bulk_load(dm, p.dm, static_cast<uint32_t>(sizeof(DM)), mbar);

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
  double* base64 = B.base_pos64 + size_t(env) * 3;
  // "flat here": the scene is the floor plane alone, or the base is out of reach of every other terrain surface (the reference
  // spawns robots over +-10 km in the stairs / ramp scenes).  Then the internal frame is re-centred on the base and the terrain
  // colliders are skipped; near the terrain the frame is the world frame.
  auto flat_at = [&](double x, double y) {
    return m.terrain_type == 0 || x > double(m.terr_bounds[0]) || x < double(m.terr_bounds[1]) || y > double(m.terr_bounds[2]) || y < double(m.terr_bounds[3]);
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

  Env<real, NCON, MAXDIM> e(m, w, reinterpret_cast<const Vert4<real>*>(p.vert), lane);
  e.hf = reinterpret_cast<const real*>(p.hf);
  e.boxes = reinterpret_cast<const DBox<real>*>(p.boxes);
  e.terrain_on = !flat;
  const unsigned env_g = unsigned(env + p.env_id_offset);
  real command[4] = {real(B.command[4 * env]), real(B.command[4 * env + 1]), real(B.command[4 * env + 2]), real(B.command[4 * env + 3])};
  float sim_time = B.sim_time[env];
  int step_count = B.step_count[env];

  // IMU.step (sensors/imu.py:110-139): measurement = truth + bias + noise, bias random walk; counter-based normals.
  // obs_row == nullptr advances the bias walk / counter only.
  auto imu_step = [&](float* obs_row) {
    unsigned tk = p.tick[env];
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

  // One pass = one "mj_step" with its env-side bookkeeping.  A reset is the same pass preceded by state sampling and the
  // lift loop; MODE_STEP with auto_reset runs a second (reset) pass for envs that just terminated, in the same warp.
  bool resetting = (MODE == MODE_RESET);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    unsigned status = 0;
    real u_late[4] = {0, 0, 0, 0};
    if (MODE != MODE_FORWARD && resetting) {
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
      for (int i = 0; i < 4; i++) u_late[i] = u[26 + i];  // consumed after the step
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
        // lift until no foot (calf-body) contact, quadruped_env.py:376-388
        bool cleared = false;
        int c_first = 0;
        if (flat) {
          // Flat floor: raising the base shifts every floor distance by exactly the lift, so after one collision pass the loop
          // reduces to a scalar recurrence on the calf-body contact distances (same iterates as re-running the collision stage).
          e.kinematics();
          e.collide_floor();
          real d = Num<real>::big, mg = 0;
          bool calf = false, boxy = false;
          if (lane < w.ncon) {
            const int info = w.c_info[lane], bdy = (info >> 8) & 0xff, g = info & 0xff;
            calf = bdy >= 2 && (bdy - 2) % 3 == 2;
            d = w.c_dist[lane]; mg = m.geom_margin[g];
            boxy = calf && m.geom_type[g] == GEOM_BOX;  // a box keeps at most 4 of its corners: not closed under lifting
          }
          if (!w.overflow && qs::ballot(boxy) == 0) {
            real lift = 0;
#pragma unroll 1
            for (int c = 0; c <= 100; c++) {
              const bool in = calf && !(d + lift > mg);
              const real pen = warp_max(in ? Num<real>::abs(d + lift) : real(0));
              if (qs::ballot(in) == 0) { cleared = true; break; }
              if (c == 100) break;
              lift += pen * real(1.1);
            }
            if (lane == 0) w.qpos[2] += lift;
            syncwarp();
            c_first = 101;
          }
        }
        e.calf_only = true;
#pragma unroll 1
        for (int c = c_first; c <= 100; c++) {
          e.kinematics();
          e.collide_floor();
          real pen = 0;
          bool any = false;
          for (int k = lane; k < w.ncon; k += 32) {
            const int bdy = (w.c_info[k] >> 8) & 0xff;
            if (bdy >= 2 && (bdy - 2) % 3 == 2) { any = true; pen = Num<real>::max(pen, Num<real>::abs(w.c_dist[k])); }
          }
          any = qs::ballot(any) != 0;
          pen = warp_max(pen);
          if (!any) { cleared = true; break; }
          if (c == 100) break;
          if (lane == 0) w.qpos[2] += pen * real(1.1);
          syncwarp();
#ifdef QS_PROF
          if (p.prof && lane == 0) { p.prof[size_t(env) * 32 + 29] = unsigned(c + 1); p.prof[size_t(env) * 32 + 30] = __float_as_uint(float(pen)); }
#endif
        }
        e.calf_only = false;
        if (!cleared) status |= 8u;
      }
      // zero ctrl / applied wrench / warm start / clock (quadruped_env.py:332-335, :394-395)
      if (lane < NV) w.warm[lane] = 0;
      if (lane < NU) w.ctrl[lane] = 0;
      if (lane < 6) w.applied[lane] = 0;
      sim_time = 0.f;
      syncwarp();
    }

    // ---- forward dynamics
    QS_MARK(1);
    e.forward_position();
    QS_MARK(2);
    typename Env<real, NCON, MAXDIM>::Flags fl = e.flags();  // contact masks depend on the collision stage only
    if (MODE == MODE_STEP && p.auto_reset && !resetting) {
      // Same-step auto-reset returns the post-reset state / observation of an env that terminates, so once the collision stage has
      // found a contact that terminates the episode (quadruped_env.py:1228-1248) the rest of this step cannot reach any output:
      // raise the flags, keep the IMU bias walk in step, and go straight to the reset pass.
      if (fl.invalid_mask != 0) {
        if (lane == 0) {
          if (p.reward) p.reward[env] = 0.f;
          if (p.terminated) p.terminated[env] = 1;
          if (p.truncated) p.truncated[env] = 0;
          if (p.sched_out) { const int idx = atomicAdd(p.sched_out_cnt + 1, 1); p.sched_out[p.num_envs + idx] = env; }
        }
        if (p.use_imu) imu_step(nullptr);
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
    if (m.has_imu) e.sensors();
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
      if (m.has_imu) e.sensors();
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
        if (lane < 6) a[AUX_OFF_IMU + lane] = m.has_imu ? float(w.sens[lane]) : 0.f;
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
    const bool terminated = fl.invalid_mask != 0 || fl.out_of_bounds;
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
      if (lane == 0) { p.episode[env] = p.episode[env] + 1; w.mu_floor = real(mu); w.mu_feet = real(mu); }
      if (lane < 6) B.qfrc_applied[size_t(env) * 6 + lane] = 0.f;
    }
    syncwarp();
    e.pack_obs(command, fl.contact_mask);
    QS_MARK(7);

    // ---- write back
    {
      bool ok = true;
      if (lane < NQ) ok = ok && isfinite(w.qpos[lane]);
      if (lane < NV) ok = ok && isfinite(w.qvel[lane]);
      if (qs::ballot(!ok) != 0) status |= 1u;
    }
    if (w.overflow) status |= 2u;
    if (e.solver_maxed) status |= 4u;
    if (lane < NQ) B.qpos[size_t(env) * NQ + lane] = (lane < 3) ? float(base64[lane]) : float(w.qpos[lane]);
    if (lane < NV) {
      B.qvel[size_t(env) * NV + lane] = float(w.qvel[lane]);
      B.qacc[size_t(env) * NV + lane] = float(w.qacc[lane]);
      B.qacc_warmstart[size_t(env) * NV + lane] = float(w.qacc[lane]);
    }
    float* obs = p.obs ? p.obs + size_t(env) * p.obs_dim : nullptr;
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
    if (obs && p.hm_rows > 0) {
      // sensors/heightmap columns: grid around the post-step base position / heading (heightmap.py:106-169)
      real qq[4] = {w.qpos[3], w.qpos[4], w.qpos[5], w.qpos[6]}, Rn[9];
      quat_normalize(qq);
      quat_to_mat(Rn, qq);
      const real ctr[3] = {w.qpos[0], w.qpos[1], w.qpos[2]};
      e.heightmap(ctr, Num<real>::atan2(Rn[3], Rn[0]), p.hm_rows, p.hm_cols, real(p.hm_dx), real(p.hm_dy), real(w.org[0]), real(w.org[1]),
                  obs + NOBS_BASE + (p.use_imu ? QS_NOBS_IMU : 0));
    }
    if (p.use_imu) imu_step(obs);
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
    if (MODE == MODE_STEP && p.sched_out && pass == 0 && lane == 0) {
      const bool heavy = w.ncon > 0 || e.solver_iter > 2;
      const int idx = atomicAdd(p.sched_out_cnt + (heavy ? 0 : 1), 1);
      p.sched_out[(heavy ? 0 : p.num_envs) + idx] = env;
    }
    // in-kernel auto-reset: the warp of an env that just terminated goes round once more as a reset pass
    if (MODE != MODE_STEP || !p.auto_reset || resetting || !terminated) break;
    resetting = true;
    given_state = false;
    e.ls_evals = 0;
    syncwarp();
  }
#ifdef QS_PROF
  if (p.prof && lane == 0) p.prof[size_t(env) * 32 + 15] = unsigned(clock64() - t_entry);
#endif
}

// HeightMap.update_height_map for every env (sensors/heightmap.py:106-169): one warp per env, rays spread over the lanes
template <typename real>
__global__ void __launch_bounds__(256) raycast_kernel(const KParams p) {
  const int lane = threadIdx.x & 31, env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const DModel<real>& m = *reinterpret_cast<const DModel<real>*>(p.dm);
  using W = WS<real, NCON_MAX, 3>;
  Env<real, NCON_MAX, 3> e(m, *reinterpret_cast<W*>(smem), reinterpret_cast<const Vert4<real>*>(p.vert), lane);  // workspace is never touched
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
}

using KernelFn = void (*)(const KParams);

template <typename real, int MAXDIM> struct Variant {
  static KernelFn fn(int mode) {
    switch (mode) {
      case MODE_STEP: return env_kernel<real, NCON_MAX, MAXDIM, MODE_STEP>;
      case MODE_RESET: return env_kernel<real, NCON_MAX, MAXDIM, MODE_RESET>;
      default: return env_kernel<real, NCON_MAX, MAXDIM, MODE_FORWARD>;
    }
  }
  static size_t ws_bytes() { return sizeof(WS<real, NCON_MAX, MAXDIM>); }
  static size_t dm_bytes() { return (sizeof(DModel<real>) + 127) & ~size_t(127); }
};

}  // namespace

struct QsHandle_ {
  QsConfig cfg{};
  int maxdim = 3;
  int obs_dim = 0;
  void* d_dm = nullptr;
  void* d_vert = nullptr;
  void* d_hf = nullptr;
  void* d_boxes = nullptr;
  KernelFn k_raycast = nullptr;
  int* d_sched = nullptr;      // 3 rotating buffers of [2N] env ids
  int* d_sched_cnt = nullptr;  // 3 x 2 counters
  int sched_phase = -1;        // -1: no valid list yet
  bool sched_enabled = true;
  unsigned* d_episode = nullptr;
  unsigned* d_tick = nullptr;
  float* d_aux = nullptr;
  // staging for the host-buffer entry point
  float* d_ctrl = nullptr; float* d_obs = nullptr; float* d_reward = nullptr; uint8_t* d_term = nullptr; uint8_t* d_trunc = nullptr;
  QsBuffers buf{};
  bool bound = false;
  KernelFn k_step = nullptr, k_reset = nullptr, k_forward = nullptr;
  unsigned* prof = nullptr;  // QS_PROF builds only
  int warps_per_cta = 8;
  size_t smem_bytes = 0;
  double timestep = 0.002;
  int64_t launches = 0;
  std::string err;
};

static thread_local std::string g_create_error;

static int fail(QsHandle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define QS_CUDA(h, call)                                                                         \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(h, 100 + int(e_), std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <typename real, int MAXDIM> static int setup_variant(QsHandle* h, const QsModel* model) {
  auto dm = std::make_unique<DModel<real>>();
  std::vector<Vert4<real>> verts;
  std::string err = build_dmodel<real>(*model, *dm, verts);
  if (!err.empty()) return fail(h, 2, err);
  QS_CUDA(h, cudaMalloc(&h->d_dm, sizeof(DModel<real>)));
  QS_CUDA(h, cudaMemcpy(h->d_dm, dm.get(), sizeof(DModel<real>), cudaMemcpyHostToDevice));
  QS_CUDA(h, cudaMalloc(&h->d_vert, sizeof(Vert4<real>) * verts.size()));
  QS_CUDA(h, cudaMemcpy(h->d_vert, verts.data(), sizeof(Vert4<real>) * verts.size(), cudaMemcpyHostToDevice));
  {
    std::vector<DBox<real>> boxes = build_boxes<real>(*model);
    std::vector<real> hf = build_hfield<real>(*model);
    QS_CUDA(h, cudaMalloc(&h->d_boxes, sizeof(DBox<real>) * boxes.size()));
    QS_CUDA(h, cudaMemcpy(h->d_boxes, boxes.data(), sizeof(DBox<real>) * boxes.size(), cudaMemcpyHostToDevice));
    QS_CUDA(h, cudaMalloc(&h->d_hf, sizeof(real) * hf.size()));
    QS_CUDA(h, cudaMemcpy(h->d_hf, hf.data(), sizeof(real) * hf.size(), cudaMemcpyHostToDevice));
  }
  h->k_raycast = raycast_kernel<real>;
  using V = Variant<real, MAXDIM>;
  h->k_step = V::fn(MODE_STEP); h->k_reset = V::fn(MODE_RESET); h->k_forward = V::fn(MODE_FORWARD);
  int dev = 0, max_smem = 0;
  QS_CUDA(h, cudaGetDevice(&dev));
  QS_CUDA(h, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int warps = LaunchCfg<real>::kMaxWarps;
#ifdef QS_PROF
  if (const char* ev = getenv("QS_WARPS_PER_CTA")) { const int v = atoi(ev); if (v >= 1 && v < warps) warps = v; }  // contention experiments
#endif
  while (warps > 1 && V::dm_bytes() + 128 + warps * V::ws_bytes() > size_t(max_smem)) warps--;
  h->warps_per_cta = warps;
  h->smem_bytes = V::dm_bytes() + 128 + warps * V::ws_bytes();
  if (h->smem_bytes > size_t(max_smem)) return fail(h, 3, "workspace does not fit in shared memory");
  for (KernelFn f : {h->k_step, h->k_reset, h->k_forward})
    QS_CUDA(h, cudaFuncSetAttribute(reinterpret_cast<const void*>(f), cudaFuncAttributeMaxDynamicSharedMemorySize, int(h->smem_bytes)));
  return 0;
}

// device-visible alias of a pinned (page-locked, mapped) host pointer, or nullptr for pageable memory
template <typename T> static T* mapped_alias(T* host) {
  if (!host) return nullptr;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, const_cast<typename std::remove_const<T>::type*>(host)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? static_cast<T*>(a.devicePointer) : nullptr;
}

extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }
int qs_model_sizeof(void) { return int(sizeof(QsModel)); }
int qs_config_sizeof(void) { return int(sizeof(QsConfig)); }
int qs_buffers_sizeof(void) { return int(sizeof(QsBuffers)); }
int qs_obs_dim(const QsConfig* cfg) { return QS_NOBS_BASE + (cfg->use_imu ? QS_NOBS_IMU : 0) + cfg->hm_rows * cfg->hm_cols * 3; }
const char* qs_last_error(QsHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int qs_max_contacts(QsHandle*) { return NCON_MAX; }
#ifdef QS_PROF
int qs_debug_set_prof(QsHandle* h, unsigned* prof) { h->prof = prof; return 0; }  // diagnostic builds only (scripts/warp_timeline.py)
#endif
int64_t qs_launch_count(QsHandle* h) { return h ? h->launches : 0; }

int qs_create(const QsModel* model, const QsConfig* cfg, QsHandle** out) {
  if (!model || !cfg || !out) return fail(nullptr, 1, "null argument");
  if (cfg->num_envs <= 0) return fail(nullptr, 1, "num_envs must be positive");
  if (cfg->hm_rows < 0 || cfg->hm_cols < 0 || cfg->hm_rows * cfg->hm_cols > 1024) return fail(nullptr, 1, "height-map grid out of range");
  QsHandle* h = new (std::nothrow) QsHandle_();
  if (!h) return fail(nullptr, 1, "out of host memory");
  h->cfg = *cfg;
  h->timestep = model->timestep;
  h->obs_dim = qs_obs_dim(cfg);
  cudaError_t ce = cudaSetDevice(cfg->device);
  if (ce != cudaSuccess) { int rc = fail(nullptr, 100 + int(ce), std::string("cudaSetDevice: ") + cudaGetErrorString(ce)); delete h; return rc; }
  h->maxdim = model_max_dim(*model) > 3 ? 6 : 3;
  int rc;
#ifdef QS_ONLY_F3  // tuning builds: only the fp32 / condim<=3 kernels are compiled
  rc = (cfg->precision == 0 && h->maxdim == 3) ? setup_variant<float, 3>(h, model) : fail(h, 4, "variant not compiled (QS_ONLY_F3)");
#else
  if (cfg->precision == 0) rc = h->maxdim == 3 ? setup_variant<float, 3>(h, model) : setup_variant<float, 6>(h, model);
  else rc = h->maxdim == 3 ? setup_variant<double, 3>(h, model) : setup_variant<double, 6>(h, model);
#endif
  if (rc == 0) {
    cudaError_t e1 = cudaMalloc(&h->d_episode, sizeof(unsigned) * cfg->num_envs), e2 = cudaMalloc(&h->d_tick, sizeof(unsigned) * cfg->num_envs);
    if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(h, 5, "cudaMalloc failed");
    else { cudaMemset(h->d_episode, 0, sizeof(unsigned) * cfg->num_envs); cudaMemset(h->d_tick, 0, sizeof(unsigned) * cfg->num_envs); }
  }
  if (rc != 0) { g_create_error = h->err; qs_destroy(h); return rc; }
  *out = h;
  return 0;
}

void qs_destroy(QsHandle* h) {
  if (!h) return;
  cudaFree(h->d_sched); cudaFree(h->d_sched_cnt);
  cudaFree(h->d_dm); cudaFree(h->d_vert); cudaFree(h->d_hf); cudaFree(h->d_boxes); cudaFree(h->d_episode); cudaFree(h->d_tick); cudaFree(h->d_aux);
  cudaFree(h->d_ctrl); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_term); cudaFree(h->d_trunc);
  delete h;
}

int qs_bind(QsHandle* h, const QsBuffers* b) {
  if (!h || !b) return fail(h, 1, "null argument");
  const void* const* ptrs = reinterpret_cast<const void* const*>(b);
  for (size_t i = 0; i < sizeof(QsBuffers) / sizeof(void*); i++)
    if (!ptrs[i]) return fail(h, 1, "QsBuffers has a null pointer");
  h->buf = *b;
  h->bound = true;
  return 0;
}

static KParams base_params(QsHandle* h) {
  KParams p{};
  p.dm = h->d_dm; p.vert = h->d_vert; p.hf = h->d_hf; p.boxes = h->d_boxes;
  p.hm_rows = h->cfg.hm_rows; p.hm_cols = h->cfg.hm_cols; p.hm_dx = float(h->cfg.hm_dx); p.hm_dy = float(h->cfg.hm_dy);
  p.num_envs = h->cfg.num_envs; p.obs_dim = h->obs_dim; p.use_imu = h->cfg.use_imu;
  p.max_iter = h->cfg.solver_max_iter > 0 ? h->cfg.solver_max_iter : (h->cfg.precision == 0 ? 50 : 100);
  p.tol = h->cfg.precision == 0 ? 1e-6f : 1e-8f;
  p.env_id_offset = h->cfg.env_id_offset;
  p.seed_lo = unsigned(h->cfg.seed & 0xffffffffu); p.seed_hi = unsigned(h->cfg.seed >> 32);
  p.imu_an = float(h->cfg.imu_accel_noise); p.imu_gn = float(h->cfg.imu_gyro_noise);
  p.imu_abr = float(h->cfg.imu_accel_bias_rate); p.imu_gbr = float(h->cfg.imu_gyro_bias_rate);
  p.b = h->buf; p.episode = h->d_episode; p.tick = h->d_tick;
  return p;
}

static int launch(QsHandle* h, KernelFn fn, const KParams& p, cudaStream_t s) {
  const int warps = h->warps_per_cta, grid = (h->cfg.num_envs + warps - 1) / warps;
  fn<<<grid, warps * 32, h->smem_bytes, s>>>(p);
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

static int step_impl(QsHandle* h, const float* ctrl, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated,
                     const QsResetOptions* auto_reset, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_step: handle not bound");
  if (!ctrl) return fail(h, 1, "qs_step: ctrl is null");
  KParams p = base_params(h);
  p.ctrl = ctrl; p.obs = obs; p.reward = reward; p.terminated = terminated; p.truncated = truncated;
  if (auto_reset) { p.auto_reset = 1; p.ro = *auto_reset; }
#ifdef QS_PROF
  p.prof = h->prof;
#endif
  if (h->sched_enabled) {
    const size_t n = size_t(h->cfg.num_envs);
    if (!h->d_sched) {
      QS_CUDA(h, cudaMalloc(&h->d_sched, 3 * 2 * n * sizeof(int)));
      QS_CUDA(h, cudaMalloc(&h->d_sched_cnt, 3 * 2 * sizeof(int)));
      QS_CUDA(h, cudaMemsetAsync(h->d_sched_cnt, 0, 3 * 2 * sizeof(int), static_cast<cudaStream_t>(stream)));
    }
    const int cur = h->sched_phase < 0 ? 0 : h->sched_phase, nxt = (cur + 1) % 3, clr = (cur + 2) % 3;
    if (h->sched_phase >= 0) { p.sched_in = h->d_sched + size_t(cur) * 2 * n; p.sched_in_cnt = h->d_sched_cnt + 2 * cur; }
    p.sched_out = h->d_sched + size_t(nxt) * 2 * n; p.sched_out_cnt = h->d_sched_cnt + 2 * nxt; p.sched_zero_cnt = h->d_sched_cnt + 2 * clr;
    h->sched_phase = nxt;
  }
  return launch(h, h->k_step, p, static_cast<cudaStream_t>(stream));
}

int qs_step(QsHandle* h, const float* ctrl, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated, void* stream) {
  return step_impl(h, ctrl, obs, reward, terminated, truncated, nullptr, stream);
}

int qs_step_autoreset(QsHandle* h, const float* ctrl, const QsResetOptions* opt, float* obs, float* reward, uint8_t* terminated,
                      uint8_t* truncated, void* stream) {
  if (!opt) return fail(h, 1, "qs_step_autoreset: options are null");
  return step_impl(h, ctrl, obs, reward, terminated, truncated, opt, stream);
}

int qs_step_host(QsHandle* h, const float* ctrl, const QsResetOptions* auto_reset, float* obs, float* reward, uint8_t* terminated,
                 uint8_t* truncated, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_step_host: handle not bound");
  if (!ctrl) return fail(h, 1, "qs_step_host: ctrl is null");
  const size_t n = size_t(h->cfg.num_envs);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->d_ctrl) {
    QS_CUDA(h, cudaMalloc(&h->d_ctrl, n * NU * sizeof(float)));
    QS_CUDA(h, cudaMalloc(&h->d_obs, n * h->obs_dim * sizeof(float)));
    QS_CUDA(h, cudaMalloc(&h->d_reward, n * sizeof(float)));
    QS_CUDA(h, cudaMalloc(&h->d_term, n));
    QS_CUDA(h, cudaMalloc(&h->d_trunc, n));
  }
  // ctrl: a pinned buffer is read by the kernel through its mapped alias (48 B per env, one coalesced row per warp, overlapped with
  // the TMA staging of the model) -- no separate H2D copy in front of the launch; pageable memory is staged with cudaMemcpyAsync.
  const float* k_ctrl = mapped_alias(ctrl);
  if (!k_ctrl) QS_CUDA(h, cudaMemcpyAsync(h->d_ctrl, ctrl, n * NU * sizeof(float), cudaMemcpyHostToDevice, s));
  // Pinned host buffers are written by the kernel itself through their mapped device alias (zero-copy): every warp streams its
  // 908-B observation row over PCIe as soon as its env is done, so the transfer overlaps with the envs still being solved instead
  // of following the kernel.  Pageable buffers fall back to staging + cudaMemcpyAsync.
  float* k_obs = mapped_alias(obs); float* k_rew = mapped_alias(reward);
  uint8_t* k_term = mapped_alias(terminated); uint8_t* k_trunc = mapped_alias(truncated);
  int rc = step_impl(h, k_ctrl ? k_ctrl : h->d_ctrl, obs ? (k_obs ? k_obs : h->d_obs) : nullptr, reward ? (k_rew ? k_rew : h->d_reward) : nullptr,
                     terminated ? (k_term ? k_term : h->d_term) : h->d_term, truncated ? (k_trunc ? k_trunc : h->d_trunc) : nullptr,
                     auto_reset, stream);
  if (rc) return rc;
  if (obs && !k_obs) QS_CUDA(h, cudaMemcpyAsync(obs, h->d_obs, n * h->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (reward && !k_rew) QS_CUDA(h, cudaMemcpyAsync(reward, h->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (terminated && !k_term) QS_CUDA(h, cudaMemcpyAsync(terminated, h->d_term, n, cudaMemcpyDeviceToHost, s));
  if (truncated && !k_trunc) QS_CUDA(h, cudaMemcpyAsync(truncated, h->d_trunc, n, cudaMemcpyDeviceToHost, s));
  QS_CUDA(h, cudaStreamSynchronize(s));
  return 0;
}

static int reset_impl(QsHandle* h, const uint8_t* mask, const float* qpos, const float* qvel, const QsResetOptions* opt, float* obs, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_reset: handle not bound");
  if (!opt) return fail(h, 1, "qs_reset: options are null");
  if ((qpos == nullptr) != (qvel == nullptr)) return fail(h, 1, "qs_reset: qpos and qvel must be given together");
  KParams p = base_params(h);
  p.mask = mask; p.in_qpos = qpos; p.in_qvel = qvel; p.ro = *opt; p.obs = obs;
  return launch(h, h->k_reset, p, static_cast<cudaStream_t>(stream));
}

int qs_reset(QsHandle* h, const uint8_t* mask, const float* qpos, const float* qvel, const QsResetOptions* opt, float* obs, void* stream) {
  return reset_impl(h, mask, qpos, qvel, opt, obs, stream);
}

int qs_reset_done(QsHandle* h, const uint8_t* terminated, const QsResetOptions* opt, float* obs, void* stream) {
  if (!terminated) return fail(h, 1, "qs_reset_done: terminated is null");
  return reset_impl(h, terminated, nullptr, nullptr, opt, obs, stream);
}

int qs_forward(QsHandle* h, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_forward: handle not bound");
  if (!h->d_aux) QS_CUDA(h, cudaMalloc(&h->d_aux, size_t(h->cfg.num_envs) * AUX_STRIDE * sizeof(float)));
  KParams p = base_params(h);
  p.aux = h->d_aux;
  return launch(h, h->k_forward, p, static_cast<cudaStream_t>(stream));
}

int qs_get(QsHandle* h, int field, float* dst, void* stream) {
  if (!h || !h->d_aux) return fail(h, 1, "qs_get: call qs_forward first");
  int off, width;
  switch (field) {
    case QS_FIELD_MASS_MATRIX: off = AUX_OFF_M; width = 324; break;
    case QS_FIELD_QFRC_BIAS: off = AUX_OFF_BIAS; width = 18; break;
    case QS_FIELD_QFRC_PASSIVE: off = AUX_OFF_PASSIVE; width = 18; break;
    case QS_FIELD_FEET_JACP: off = AUX_OFF_JACP; width = 216; break;
    case QS_FIELD_FEET_POS: off = AUX_OFF_FEETPOS; width = 12; break;
    case QS_FIELD_COM: off = AUX_OFF_COM; width = 3; break;
    case QS_FIELD_CONTACTS: off = AUX_OFF_CONTACTS; width = QS_CONTACT_STRIDE * NCON_MAX; break;
    case QS_FIELD_QFRC_SMOOTH: off = AUX_OFF_SMOOTH; width = 18; break;
    case QS_FIELD_QFRC_CONSTRAINT: off = AUX_OFF_CONSTRAINT; width = 18; break;
    case QS_FIELD_XPOS: off = AUX_OFF_XPOS; width = 39; break;
    case QS_FIELD_SENSOR_IMU: off = AUX_OFF_IMU; width = 6; break;
    case QS_FIELD_FEET_JACR: off = AUX_OFF_JACR; width = 216; break;
    case QS_FIELD_FEET_JACP_DOT: off = AUX_OFF_JACP_DOT; width = 216; break;
    case QS_FIELD_FEET_JACR_DOT: off = AUX_OFF_JACR_DOT; width = 216; break;
    default: return fail(h, 1, "qs_get: unknown field");
  }
  QS_CUDA(h, cudaMemcpy2DAsync(dst, width * sizeof(float), h->d_aux + off, AUX_STRIDE * sizeof(float), width * sizeof(float), h->cfg.num_envs,
                               cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

int qs_raycast_heightmap(QsHandle* h, int rows, int cols, double dx, double dy, float* out, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_raycast_heightmap: handle not bound");
  if (rows <= 0 || cols <= 0 || !out) return fail(h, 1, "qs_raycast_heightmap: bad arguments");
  KParams p = base_params(h);
  p.hm_rows = rows; p.hm_cols = cols; p.hm_dx = float(dx); p.hm_dy = float(dy); p.hm_out = out;
  const int warps = 8, grid = (h->cfg.num_envs + warps - 1) / warps;
  h->k_raycast<<<grid, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
