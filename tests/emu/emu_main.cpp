// emu_main.cpp -- TEST INFRASTRUCTURE. Host build of the fused-step device code (see warp_emu.h).
// Exposes one C entry point that advances a single environment by one step with 32 fibers acting as the warp.
#include "warp_emu.h"

#include <memory>
#include <vector>

#include "../../gym_quadruped_b200/csrc/qs_host_model.h"

namespace {
template <typename real, int MAXDIM, int FEAT>
int run_step(const QsModel* model, double* qpos, double* qvel, double* warm, const double* ctrl, const double* envp, double* obs,
             double* misc, int max_iter, double tol, int mode) {
  using namespace qs;
  constexpr int NCON = 16;
  auto dm = std::make_unique<DModel<real>>();
  std::vector<Vert4<real>> verts;
  std::string err = build_dmodel<real>(*model, *dm, verts);
  if (!err.empty()) return -1;
  std::vector<DBox<real>> boxes = build_boxes<real>(*model);
  std::vector<real> hfv = build_hfield<real>(*model);
  auto ws = std::make_unique<WS<real, NCON, MAXDIM>>();
  std::memset(ws.get(), 0, sizeof(*ws));
  const bool flat = model->terrain_type == QS_TERRAIN_FLAT;
  ws->org[0] = flat ? std::nearbyint(qpos[0]) : 0.0;
  ws->org[1] = flat ? std::nearbyint(qpos[1]) : 0.0;
  for (int i = 0; i < 19; i++) ws->qpos[i] = real(i < 2 ? qpos[i] - ws->org[i] : qpos[i]);
  for (int i = 0; i < 18; i++) { ws->qvel[i] = real(qvel[i]); ws->warm[i] = real(warm[i]); }
  for (int i = 0; i < 12; i++) ws->ctrl[i] = real(ctrl[i]);
  ws->mu_floor = real(envp[0]); ws->mu_feet = real(envp[1]);
  real command[4];
  for (int i = 0; i < 4; i++) command[i] = real(envp[2 + i]);
  for (int i = 0; i < 6; i++) ws->applied[i] = real(envp[6 + i]);
  double base64[3] = {qpos[0], qpos[1], qpos[2]};
  WarpCtx ctx;
  int iters = 0, maxed = 0;
  std::vector<float> bias_buf(18, 0.f);
  unsigned cmask = 0, imask = 0;
  static float hm_out[75];
  bool oob = false;
  {
    ctx.run([&](int lane) {
      Env<real, NCON, MAXDIM, FEAT> e(*dm, *ws, verts.data(), lane);
      e.bias_out = bias_buf.data();
      e.hf = hfv.data(); e.boxes = boxes.data();
      e.forward(max_iter, real(tol));
      auto f = e.flags();
      if (mode == 1) {
        e.integrate(base64);
        f.out_of_bounds = e.out_of_bounds();
        e.pack_obs(command, f.contact_mask);
      }
      if (mode == 2) {  // height map around the current base position / heading
        const real* q = ws->qpos;
        real R[9]; real qq[4] = {q[3], q[4], q[5], q[6]}; quat_normalize(qq); quat_to_mat(R, qq);
        real ctr[3] = {q[0], q[1], q[2]};
        e.heightmap(ctr, Num<real>::atan2(R[3], R[0]), 5, 5, real(0.1), real(0.1), real(ws->org[0]), real(ws->org[1]), hm_out);
      }
      if (lane == 0) { iters = e.solver_iter; maxed = e.solver_maxed; cmask = f.contact_mask; imask = f.invalid_mask; oob = f.out_of_bounds; }
    });
  }
  // results
  for (int i = 0; i < 19; i++) qpos[i] = double(ws->qpos[i]) + (i < 2 ? ws->org[i] : 0.0);
  if (mode == 1) for (int i = 0; i < 3; i++) qpos[i] = base64[i];
  for (int i = 0; i < 18; i++) { qvel[i] = double(ws->qvel[i]); warm[i] = double(ws->qacc[i]); }
  if (obs) for (int i = 0; i < 227; i++) obs[i] = double(ws->obs[i]);
  if (misc) {
    int k = 0;
    misc[k++] = iters; misc[k++] = maxed; misc[k++] = cmask; misc[k++] = imask; misc[k++] = oob; misc[k++] = ws->ncon; misc[k++] = ws->overflow;
    misc[k++] = FEAT;
    k = 8;
    for (int i = 0; i < 18; i++) misc[k++] = double(ws->qacc[i]);      // 8
    for (int i = 0; i < 18; i++) misc[k++] = double(bias_buf[i]);      // 26
    for (int i = 0; i < 18; i++) misc[k++] = double(ws->fsm[i]);       // 44
    for (int i = 0; i < 18; i++) misc[k++] = double(ws->asmooth[i]);   // 62
    for (int i = 0; i < 18; i++) misc[k++] = double(ws->fcon[i]);      // 80
    // dense M, 98
    for (int i = 0; i < 18; i++)
      for (int j = 0; j < 18; j++) {
        double v = 0;
        if (i < 6 && j < 6) v = ws->Mbb[i][j];
        else if (i >= 6 && j < 6) v = ws->Mlb[(i - 6) / 3][(i - 6) % 3][j];
        else if (i < 6 && j >= 6) v = ws->Mlb[(j - 6) / 3][(j - 6) % 3][i];
        else if ((i - 6) / 3 == (j - 6) / 3) v = ws->Mll[(i - 6) / 3][(i - 6) % 3][(j - 6) % 3];
        misc[k++] = v;
      }
    // contacts, 422: per contact dist,pos3,frame9,F3,geom,body,mu,dim
    for (int c = 0; c < ws->ncon; c++) {
      const int info = ws->c_info[c], dim = (info >> 16) & 0xff;
      misc[k++] = ws->c_dist[c];
      misc[k++] = ws->c_pos[c][0] + ws->org[0]; misc[k++] = ws->c_pos[c][1] + ws->org[1]; misc[k++] = ws->c_pos[c][2];
      const real* f = ws->c_frame[c];
      for (int i = 0; i < 6; i++) misc[k++] = f[i];
      misc[k++] = f[1] * f[5] - f[2] * f[4]; misc[k++] = f[2] * f[3] - f[0] * f[5]; misc[k++] = f[0] * f[4] - f[1] * f[3];
      for (int i = 0; i < 3; i++) misc[k++] = (i < dim) ? double(ws->c_F[c][i]) : 0.0;
      misc[k++] = info & 0xff; misc[k++] = (info >> 8) & 0xff; misc[k++] = ws->c_fri[c][0]; misc[k++] = dim;
    }
    // sensors at 422 + 20*16 = 742
    k = 742;
    for (int i = 0; i < 6; i++) misc[k++] = double(ws->sens[i]);
    for (int i = 0; i < 75; i++) misc[k++] = double(hm_out[i]);  // 748
  }
  return 0;
}
}  // namespace

// mode 0: forward only; mode 1: forward + integrate + obs; mode 2: + height map. precision 0: fp32, 1: fp64.
// envp = {mu_floor, mu_feet, command[4], applied[6]}; misc has room for 1024 doubles.
// feat_sel 0: the generic variant (FEAT = 0); 1: the specialised variant the library would pick for this model (csrc/qs_variants.h:
// the four BASELINE configurations, chosen by the same `model_features` rule as qstep.cu), generic if none applies.
extern "C" int emu_step(const QsModel* model, int precision, double* qpos, double* qvel, double* warm, const double* ctrl,
                        const double* envp, double* obs, double* misc, int max_iter, double tol, int mode, int feat_sel) {
  using namespace qs;
  const int md = model_max_dim(*model);
#define QS_EMU_RUN(real, MD, F) run_step<real, MD, F>(model, qpos, qvel, warm, ctrl, envp, obs, misc, max_iter, tol, mode)
  if (feat_sel == 1) {
    // as the library with use_imu = 0: IMU columns are not part of the emulated observation row, so a variant that drops the sensor
    // stage is admissible (the `imu` entries of misc are then not filled); mode 2 asks for the height map
    const int have = model_features(*model, false, mode == 2 ? 25 : 0);
    auto ok = [&](int feat) { return (feat & ~have) == 0; };
    if (precision == 0) {
      if (md <= 3 && ok(FEAT_CFG2)) return QS_EMU_RUN(float, 3, FEAT_CFG2);
      if (md <= 3 && ok(FEAT_CFG3)) return QS_EMU_RUN(float, 3, FEAT_CFG3);
      if (md <= 3 && ok(FEAT_CFG5)) return QS_EMU_RUN(float, 3, FEAT_CFG5);
      if (md > 3 && ok(FEAT_CFG4)) return QS_EMU_RUN(float, 6, FEAT_CFG4);
      if (md <= 3 && ok(FEAT_PYR_FLAT_PRIM)) return QS_EMU_RUN(float, 3, FEAT_PYR_FLAT_PRIM);
      if (md > 3 && ok(FEAT_ELL_FLAT_MESH6)) return QS_EMU_RUN(float, 6, FEAT_ELL_FLAT_MESH6);
      if (md > 3 && ok(FEAT_ELL_FLAT_PRIM)) return QS_EMU_RUN(float, 6, FEAT_ELL_FLAT_PRIM);
    } else {
      if (md <= 3 && ok(FEAT_CFG2)) return QS_EMU_RUN(double, 3, FEAT_CFG2);
      if (md <= 3 && ok(FEAT_CFG3)) return QS_EMU_RUN(double, 3, FEAT_CFG3);
      if (md <= 3 && ok(FEAT_CFG5)) return QS_EMU_RUN(double, 3, FEAT_CFG5);
      if (md > 3 && ok(FEAT_CFG4)) return QS_EMU_RUN(double, 6, FEAT_CFG4);
      if (md <= 3 && ok(FEAT_PYR_FLAT_PRIM)) return QS_EMU_RUN(double, 3, FEAT_PYR_FLAT_PRIM);
      if (md > 3 && ok(FEAT_ELL_FLAT_MESH6)) return QS_EMU_RUN(double, 6, FEAT_ELL_FLAT_MESH6);
      if (md > 3 && ok(FEAT_ELL_FLAT_PRIM)) return QS_EMU_RUN(double, 6, FEAT_ELL_FLAT_PRIM);
    }
  }
  if (precision == 0) return md > 3 ? QS_EMU_RUN(float, 6, 0) : QS_EMU_RUN(float, 3, 0);
  return md > 3 ? QS_EMU_RUN(double, 6, 0) : QS_EMU_RUN(double, 3, 0);
#undef QS_EMU_RUN
}
