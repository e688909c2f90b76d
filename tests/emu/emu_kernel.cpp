// emu_kernel.cpp -- TEST INFRASTRUCTURE.  Runs the step / reset / forward KERNEL (gym_quadruped_b200/csrc/qs_kernel.cuh, generic
// variant) on the host warp emulator, one emulated CTA of one warp per environment, against caller-owned host buffers laid out
// exactly like the device buffers of the C-ABI (QsBuffers).  The parameter block is filled the way qstep.cu does it
// (base_params / step_impl / reset_impl, plain stream order: q_sync = 0).
#define QS_HOST_EMU
#include "kernel_emu.h"

#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../gym_quadruped_b200/csrc/qs_host_model.h"
#include "../../gym_quadruped_b200/csrc/qs_kernel.cuh"

extern "C" {
struct EmuLaunch {
  int precision, mode, num_envs, use_imu, hm_rows, hm_cols;
  double hm_dx, hm_dy;
  int max_iter, env_id_offset, auto_reset, pad0;
  double tol;
  uint64_t seed;
  double imu_an, imu_gn, imu_abr, imu_gbr;
  QsBuffers buf;
  unsigned *episode, *tick, *cmd_epoch, *ext_epoch;
  QsSchedule sched;
  const float* ctrl;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  const uint8_t* mask;
  const float* in_qpos;
  const float* in_qvel;
  QsResetOptions ro;
  float* aux;
  // fused observation gather (qs_gather_*): `gather_world` host tensors stand in for the peer-mapped ones
  int gather_world, gather_rank, gather_row_stride, gather_seq;
  float* gather_peers[8];
  unsigned* gather_flags[8];
};
int emu_launch_sizeof(void) { return int(sizeof(EmuLaunch)); }
int emu_aux_stride(void) { return qs::AUX_STRIDE; }
}

namespace {
using namespace qs;

template <typename real, int MAXDIM>
int run_kernel(const QsModel* model, const EmuLaunch& L) {
  constexpr int NCON = NCON_MAX;
  auto dm = std::make_unique<DModel<real>>();
  std::vector<Vert4<real>> verts;
  if (!build_dmodel<real>(*model, *dm, verts).empty()) return -1;
  std::vector<DBox<real>> boxes = build_boxes<real>(*model);
  std::vector<real> hf = build_hfield<real>(*model);
  const int n = L.num_envs;
  KParams p{};
  p.dm = dm.get(); p.vert = verts.data(); p.hf = hf.data(); p.boxes = boxes.data();
  p.hm_rows = L.hm_rows; p.hm_cols = L.hm_cols; p.hm_dx = float(L.hm_dx); p.hm_dy = float(L.hm_dy);
  p.num_envs = n; p.obs_dim = QS_NOBS_BASE + (L.use_imu ? QS_NOBS_IMU : 0) + L.hm_rows * L.hm_cols * 3; p.obs_stride = p.obs_dim;
  p.use_imu = L.use_imu; p.max_iter = L.max_iter; p.tol = float(L.tol); p.env_id_offset = L.env_id_offset;
  p.seed_lo = unsigned(L.seed & 0xffffffffu); p.seed_hi = unsigned(L.seed >> 32);
  p.sch_command_mode = L.sched.command_mode; p.sch_ext_enabled = L.sched.ext_enabled;
  for (int i = 0; i < 2; i++) { p.sch_lin[i] = float(L.sched.lin_vel_range[i]); p.sch_ang[i] = float(L.sched.ang_vel_range[i]); }
  for (int i = 0; i < 6; i++) { p.sch_ext_lo[i] = float(L.sched.ext_lo[i]); p.sch_ext_hi[i] = float(L.sched.ext_hi[i]); }
  p.cmd_epoch = L.cmd_epoch; p.ext_epoch = L.ext_epoch;
  p.imu_an = float(L.imu_an); p.imu_gn = float(L.imu_gn); p.imu_abr = float(L.imu_abr); p.imu_gbr = float(L.imu_gbr);
  p.b = L.buf; p.episode = L.episode; p.tick = L.tick;
  p.ctrl = L.ctrl; p.obs = L.obs; p.reward = L.reward; p.terminated = L.terminated; p.truncated = L.truncated;
  p.mask = L.mask; p.in_qpos = L.in_qpos; p.in_qvel = L.in_qvel; p.ro = L.ro; p.aux = L.aux;
  p.auto_reset = L.auto_reset;
  if (L.gather_world > 1) {  // as in step_impl: own rows live in this rank's block of its own gathered tensor
    p.gather_world = L.gather_world; p.gather_rank = L.gather_rank; p.gather_seq = unsigned(L.gather_seq);
    for (int q = 0; q < L.gather_world; q++) { p.gather_peers[q] = L.gather_peers[q]; p.gather_flags[q] = L.gather_flags[q]; }
    p.obs = p.gather_peers[L.gather_rank] + size_t(L.gather_rank) * n * L.gather_row_stride;
    p.obs_stride = L.gather_row_stride;
  }
  // finish-order queues in plain stream order: identity placement, generation 0
  std::vector<int> q_in(n), q_out(n, -1 - QS_SLOT_GEN_MASK);
  for (int i = 0; i < n; i++) q_in[i] = i;
  unsigned tail = 0;
  p.q_in = q_in.data(); p.q_out = q_out.data(); p.q_tail = &tail; p.q_tail_base = 0; p.q_contiguous = 1; p.q_sync = 0;
  p.q_gen_in = 0; p.q_gen_out = 0;

  using W = WS<real, NCON, MAXDIM>;
  constexpr size_t DM_BYTES = (sizeof(DModel<real>) + 127) & ~size_t(127);
  const size_t smem_bytes = DM_BYTES + 128 + sizeof(W);
  unsigned char* smem = static_cast<unsigned char*>(std::aligned_alloc(128, (smem_bytes + 127) & ~size_t(127)));
  for (int env = 0; env < n; env++) {
    std::memset(smem, 0, smem_bytes);
    WarpCtx ctx;
    blockIdx.x = unsigned(env); gridDim.x = unsigned(n); blockDim.x = 32;
    g_smem = smem;
    ctx.run([&](int) {
      if (L.mode == 0) env_kernel<real, NCON, MAXDIM, MODE_STEP, 0>(p);
      else if (L.mode == 1) env_kernel<real, NCON, MAXDIM, MODE_RESET, 0>(p);
      else env_kernel<real, NCON, MAXDIM, MODE_FORWARD, 0>(p);
    });
  }
  std::free(smem);
  for (int i = 0; i < n; i++)
    if (L.mode == 0 && (q_out[i] < 0 || (q_out[i] & ((1 << QS_SLOT_ENV_BITS) - 1)) >= n)) return -2;  // every env published once
  return 0;
}
}  // namespace

// One launch of the generic kernel variant over `num_envs` environments.  mode 0 step (auto_reset optional), 1 reset, 2 forward.
extern "C" int emu_kernel(const QsModel* model, const EmuLaunch* L) {
  const int md = qs::model_max_dim(*model);
  if (L->precision == 0) return md > 3 ? run_kernel<float, 6>(model, *L) : run_kernel<float, 3>(model, *L);
  return md > 3 ? run_kernel<double, 6>(model, *L) : run_kernel<double, 3>(model, *L);
}
