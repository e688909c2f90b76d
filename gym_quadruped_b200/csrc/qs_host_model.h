// qs_host_model.h -- host-side conversion of the C-ABI `QsModel` (fp64 tables, include/qstep.h) into the kernel's
// `DModel<real>`: precision cast plus every quantity that is constant per robot+scene and would otherwise be recomputed
// per contact per step (contact-parameter mixing against the terrain geoms, solref -> (K, B), friction-loss R/D).
// Follows the engine's mj_contactParam / getsolparam rules.  Plain C++ (no CUDA), so the host
// warp emulator under tests/emu uses the very same code.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/qstep.h"
#include "qs_env.cuh"

namespace qs {

inline void solparam(const double* solref, const double* solimp, double timestep, double* K, double* B) {
  const double dmax = std::min(std::max(solimp[1], 0.0001), 0.9999);
  const double tc = std::max(solref[0], 2 * timestep), dr = solref[1];
  *K = 1.0 / std::max(1e-15, dmax * dmax * tc * tc * dr * dr);
  *B = 2.0 / std::max(1e-15, dmax * tc);
}

inline void quat2mat(const double* q, double* m) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

// returns "" on success, else an error message
template <typename real>
std::string build_dmodel(const QsModel& s, DModel<real>& d, std::vector<Vert4<real>>& verts) {
  if (s.abi_version != QS_ABI_VERSION) return "QsModel.abi_version mismatch";
  if (s.ngeom < 0 || s.ngeom > QS_MAXGEOM) return "ngeom out of range";
  std::fill_n(reinterpret_cast<unsigned char*>(&d), sizeof(d), 0);
  const double big = sizeof(real) == 4 ? 1e30 : 1e300;
  d.timestep = real(s.timestep);
  for (int i = 0; i < 3; i++) d.gravity[i] = real(s.gravity[i]);
  d.impratio = real(s.impratio); d.tolerance = real(s.tolerance); d.ls_tolerance = real(s.ls_tolerance); d.meaninertia = real(s.meaninertia);
  for (int i = 0; i < 4; i++) d.terrain_limits[i] = real(s.terrain_limits[i]);
  for (int i = 0; i < 3; i++) d.floor_fri[i] = real(s.floor_par.friction[i]);
  for (int i = 0; i < 3; i++) d.imu_pos[i] = real(s.imu_pos[i]);
  { double m9[9]; quat2mat(s.imu_quat, m9); for (int i = 0; i < 9; i++) d.imu_mat[i] = real(m9[i]); }
  double mass = 0;
  for (int b = 0; b < QS_NBODY; b++) {
    mass += s.body_mass[b];
    for (int i = 0; i < 3; i++) { d.body_pos[b][i] = real(s.body_pos[b][i]); d.body_ipos[b][i] = real(s.body_ipos[b][i]); d.body_inertia[b][i] = real(s.body_inertia[b][i]); }
    for (int i = 0; i < 4; i++) d.body_quat[b][i] = real(s.body_quat[b][i]);
    { double m9[9]; quat2mat(s.body_iquat[b], m9); for (int i = 0; i < 9; i++) d.body_imat[b][i] = real(m9[i]); }
    d.body_mass[b] = real(s.body_mass[b]);
    d.body_iw[b][0] = real(s.body_invweight0[b][0]); d.body_iw[b][1] = real(s.body_invweight0[b][1]);
    const int expect = b <= 1 ? 0 : ((b - 2) % 3 == 0 ? 1 : b - 1);
    if (s.body_parent[b] != expect) return "body tree is not base + 4 x (hip, thigh, calf)";
  }
  d.mass_total = real(mass);
  {  // conservative reach of any robot geom from the base origin (terrain broad phase)
    double reach = 0;
    for (int l = 0; l < 4; l++) {
      double chain = 0;
      for (int k = 0; k < 3; k++) { const double* bp = s.body_pos[2 + 3 * l + k]; chain += std::sqrt(bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]); }
      reach = std::max(reach, chain);
    }
    double gext = 0;
    for (int g = 0; g < s.ngeom; g++) { const double* gp = s.geom_pos[g]; gext = std::max(gext, std::sqrt(gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]) + s.geom_rbound[g]); }
    d.robot_radius = real(reach + gext + 0.05);
  }
  for (int i = 0; i < 4; i++) d.hf_size[i] = real(s.hf_size[i]);
  for (int i = 0; i < 3; i++) { d.hf_pos[i] = real(s.hf_pos[i]); d.terr_fri[i] = real(s.box_par.friction[i]); }
  d.terr_margin = real(s.box_par.margin);
  {  // boxes that out-rank every robot geom (scene_slippery.xml: priority 2) impose their own contact parameters [MJ mj_contactParam]
    int max_prio = -1000000;
    for (int g = 0; g < s.ngeom; g++) max_prio = std::max(max_prio, int(s.geom_par[g].priority));
    d.terr_wins = (s.terrain_type == QS_TERRAIN_BOXES && s.box_par.priority > max_prio) ? 1 : 0;
    if (s.terrain_type == QS_TERRAIN_BOXES && s.box_par.priority != 0 && !d.terr_wins) return "box priority must be 0 or above every robot geom";
    if (s.box_par.margin != s.floor_par.margin || s.box_par.gap != s.floor_par.gap) return "terrain boxes must share the floor's margin / gap";
    double K = 0, B = 0;
    if (d.terr_wins) {
      if (s.box_par.solref[0] <= 0) return "direct (negative) solref not supported";
      solparam(s.box_par.solref, s.box_par.solimp, s.timestep, &K, &B);
      if (!(s.box_par.condim == 1 || s.box_par.condim == 3)) return "unsupported box contact dimensionality";
    }
    d.terr_K = real(K); d.terr_B = real(B); d.terr_dim = s.box_par.condim;
    for (int i = 0; i < 5; i++) d.terr_solimp[i] = real(s.box_par.solimp[i]);
  }
  {  // bounding rectangle of the non-floor terrain, padded by the robot's reach: outside it an env sees the floor plane only
    double b[4] = {-big, big, -big, big};
    const double pad = double(d.robot_radius) + 1.0;
    if (s.terrain_type == QS_TERRAIN_HFIELD) {
      b[0] = s.hf_pos[0] + s.hf_size[0]; b[1] = s.hf_pos[0] - s.hf_size[0]; b[2] = s.hf_pos[1] + s.hf_size[1]; b[3] = s.hf_pos[1] - s.hf_size[1];
    } else if (s.terrain_type == QS_TERRAIN_BOXES) {
      for (int k = 0; k < s.nbox; k++) {
        const double* h = s.box_half[k];
        const double r = std::sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
        b[0] = std::max(b[0], s.box_pos[k][0] + r); b[1] = std::min(b[1], s.box_pos[k][0] - r);
        b[2] = std::max(b[2], s.box_pos[k][1] + r); b[3] = std::min(b[3], s.box_pos[k][1] - r);
      }
    }
    d.terr_bounds[0] = real(b[0] + pad); d.terr_bounds[1] = real(b[1] - pad); d.terr_bounds[2] = real(b[2] + pad); d.terr_bounds[3] = real(b[3] - pad);
  }
  d.hf_nrow = s.hf_nrow; d.hf_ncol = s.hf_ncol;
  if (s.terrain_type == QS_TERRAIN_HFIELD && (s.hf_nrow < 2 || s.hf_ncol < 2 || !s.hf_data)) return "height field data missing";
  if (s.terrain_type == QS_TERRAIN_HFIELD) {  // largest gradient magnitude over both triangles of every cell (+ a rounding cushion)
    const int nc = s.hf_ncol, nr = s.hf_nrow;
    const double dx = 2 * s.hf_size[0] / (nc - 1), dy = 2 * s.hf_size[1] / (nr - 1), sz = s.hf_size[2];
    double lip2 = 0;
    for (int r = 0; r + 1 < nr; r++)
      for (int c = 0; c + 1 < nc; c++) {
        const double z00 = sz * s.hf_data[r * nc + c], z10 = sz * s.hf_data[r * nc + c + 1], z01 = sz * s.hf_data[(r + 1) * nc + c], z11 = sz * s.hf_data[(r + 1) * nc + c + 1];
        const double g1x = (z10 - z00) / dx, g1y = (z11 - z10) / dy, g2x = (z11 - z01) / dx, g2y = (z01 - z00) / dy;
        lip2 = std::max(lip2, std::max(g1x * g1x + g1y * g1y, g2x * g2x + g2y * g2y));
      }
    d.hf_lip = real(std::sqrt(lip2) * 1.001 + 1e-6);
  }
  if (s.terrain_type == QS_TERRAIN_BOXES && (s.nbox < 0 || s.nbox > QS_MAXBOX)) return "nbox out of range";
  for (int j = 0; j < QS_NJNT; j++) {
    for (int i = 0; i < 3; i++) {
      if (s.jnt_pos[j][i] != 0.0) return "joint anchors offset from the body origin (jnt_pos != 0) are not supported";
      d.jnt_pos[j][i] = real(s.jnt_pos[j][i]); d.jnt_axis[j][i] = real(s.jnt_axis[j][i]);
    }
    d.jnt_range[j][0] = real(s.jnt_range[j][0]); d.jnt_range[j][1] = real(s.jnt_range[j][1]);
    double K, B;
    if (s.jnt_solref[j][0] <= 0) return "direct (negative) solref not supported";
    solparam(s.jnt_solref[j], s.jnt_solimp[j], s.timestep, &K, &B);
    d.jnt_K[j] = real(K); d.jnt_B[j] = real(B);
    for (int i = 0; i < 5; i++) d.jnt_solimp[j][i] = real(s.jnt_solimp[j][i]);
    d.jnt_margin[j] = real(s.jnt_margin[j]);
    d.jnt_limited[j] = s.jnt_limited[j];
  }
  for (int i = 0; i < QS_NQ; i++) { d.qpos0[i] = real(s.qpos0[i]); d.key_qpos[i] = real(s.key_qpos[i]); }
  for (int k = 0; k < QS_NV; k++) {
    d.dof_damping[k] = real(s.dof_damping[k]); d.dof_armature[k] = real(s.dof_armature[k]); d.dof_floss[k] = real(s.dof_frictionloss[k]);
    d.dof_iw[k] = real(s.dof_invweight0[k]);
    double K, B;
    if (s.dof_solref[k][0] <= 0) return "direct (negative) solref not supported";
    solparam(s.dof_solref[k], s.dof_solimp[k], s.timestep, &K, &B);
    d.dof_B[k] = real(B);
    // friction-loss rows sit at pos - margin = 0 -> impedance d0
    const double d0 = std::min(std::max(s.dof_solimp[k][0], 0.0001), 0.9999), d1 = std::min(std::max(s.dof_solimp[k][1], 0.0001), 0.9999);
    const double imp = (d0 == d1 || s.dof_solimp[k][2] <= 1e-15) ? 0.5 * (d0 + d1) : d0;
    const double R = std::max(1e-15, (1 - imp) * s.dof_invweight0[k] / imp);
    d.dof_R[k] = real(R); d.dof_D[k] = real(1.0 / R);
    if (k < 6 && s.dof_frictionloss[k] != 0) return "friction loss on the free joint not supported";
  }
  for (int a = 0; a < QS_NU; a++) {
    d.act_clo[a] = real(s.act_ctrllimited[a] ? s.act_ctrlrange[a][0] : -big); d.act_chi[a] = real(s.act_ctrllimited[a] ? s.act_ctrlrange[a][1] : big);
    d.act_flo[a] = real(s.act_forcelimited[a] ? s.act_forcerange[a][0] : -big); d.act_fhi[a] = real(s.act_forcelimited[a] ? s.act_forcerange[a][1] : big);
  }
  d.cone = s.cone; d.iterations = s.iterations; d.ls_iterations = s.ls_iterations; d.ngeom = s.ngeom; d.nvert = s.nvert;
  d.terrain_type = s.terrain_type; d.nbox = s.nbox; d.has_imu = s.has_imu;
  const QsGeomParams& wp = s.floor_par;  // every terrain geom (floor, hfield, boxes) carries default parameters
  for (int g = 0; g < s.ngeom; g++) {
    const QsGeomParams& gp = s.geom_par[g];
    const int t = s.geom_type[g];
    if (!(t == QS_GEOM_SPHERE || t == QS_GEOM_CAPSULE || t == QS_GEOM_CYLINDER || t == QS_GEOM_BOX || t == QS_GEOM_MESH)) return "unsupported geom type";
    d.geom_type[g] = t; d.geom_body[g] = s.geom_body[g]; d.geom_leg[g] = s.geom_foot_leg[g];
    d.geom_vertadr[g] = s.geom_vertadr[g]; d.geom_vertnum[g] = s.geom_vertnum[g];
    double m9[9];
    quat2mat(s.geom_quat[g], m9);
    for (int i = 0; i < 9; i++) d.geom_mat[g][i] = real(m9[i]);
    for (int i = 0; i < 3; i++) {
      d.geom_pos[g][i] = real(s.geom_pos[g][i]); d.geom_size[g][i] = real(s.geom_size[g][i]); d.geom_bcenter[g][i] = real(s.geom_bcenter[g][i]);
      d.geom_fri[g][i] = real(gp.friction[i]);
    }
    d.geom_rbound[g] = real(s.geom_rbound[g]);
    if (t == QS_GEOM_MESH) {  // body-frame bounding box of the hull (broad phase)
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int v = s.geom_vertadr[g]; v < s.geom_vertadr[g] + s.geom_vertnum[g]; v++)
        for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], s.vert[3 * v + i]); hi[i] = std::max(hi[i], s.vert[3 * v + i]); }
      for (int i = 0; i < 3; i++) { d.geom_bcenter[g][i] = real(0.5 * (lo[i] + hi[i])); d.geom_bhalf[g][i] = real(0.5 * (hi[i] - lo[i]) * (1 + 1e-6) + 1e-9); }
    }
    // mj_contactParam against a default-parameter world geom
    double solref[2], solimp[5];
    int dim;
    if (gp.priority == wp.priority) {
      dim = std::max(gp.condim, wp.condim);
      double mix;
      if (wp.solmix >= 1e-15 && gp.solmix >= 1e-15) mix = wp.solmix / (wp.solmix + gp.solmix);
      else if (wp.solmix < 1e-15 && gp.solmix < 1e-15) mix = 0.5;
      else mix = wp.solmix < 1e-15 ? 0.0 : 1.0;
      if (wp.solref[0] > 0 && gp.solref[0] > 0) for (int i = 0; i < 2; i++) solref[i] = mix * wp.solref[i] + (1 - mix) * gp.solref[i];
      else for (int i = 0; i < 2; i++) solref[i] = std::min(wp.solref[i], gp.solref[i]);
      for (int i = 0; i < 5; i++) solimp[i] = mix * wp.solimp[i] + (1 - mix) * gp.solimp[i];
      d.geom_prio[g] = 0;
    } else {
      const QsGeomParams& win = gp.priority > wp.priority ? gp : wp;
      dim = win.condim;
      for (int i = 0; i < 2; i++) solref[i] = win.solref[i];
      for (int i = 0; i < 5; i++) solimp[i] = win.solimp[i];
      d.geom_prio[g] = gp.priority > wp.priority ? 1 : -1;
    }
    if (solref[0] <= 0) return "direct (negative) solref not supported";
    if (!(dim == 1 || dim == 3 || (dim == 6 && s.cone == QS_CONE_ELLIPTIC))) return "unsupported contact dimensionality";
    d.geom_dim[g] = dim;
    double K, B;
    solparam(solref, solimp, s.timestep, &K, &B);
    d.geom_K[g] = real(K); d.geom_B[g] = real(B);
    for (int i = 0; i < 5; i++) d.geom_solimp[g][i] = real(solimp[i]);
    const double margin = std::max(gp.margin, wp.margin), gap = std::max(gp.gap, wp.gap);
    d.geom_margin[g] = real(margin); d.geom_incmargin[g] = real(margin - gap);
  }
  for (int l = 0; l < 4; l++) d.foot_geom[l] = s.foot_geom[l];
  verts.resize(std::max(1, s.nvert));
  for (int i = 0; i < s.nvert; i++) { verts[i].x = real(s.vert[3 * i]); verts[i].y = real(s.vert[3 * i + 1]); verts[i].z = real(s.vert[3 * i + 2]); verts[i].w = 0; }
  return "";
}

template <typename real> std::vector<DBox<real>> build_boxes(const QsModel& s) {
  std::vector<DBox<real>> out(std::max(1, s.nbox));
  for (int b = 0; b < s.nbox; b++) {
    double m9[9];
    quat2mat(s.box_quat[b], m9);
    double r2 = 0;
    for (int i = 0; i < 3; i++) { out[b].pos[i] = real(s.box_pos[b][i]); out[b].half[i] = real(s.box_half[b][i]); r2 += s.box_half[b][i] * s.box_half[b][i]; }
    for (int i = 0; i < 9; i++) out[b].mat[i] = real(m9[i]);
    out[b].rad = real(std::sqrt(r2));
    for (int i = 0; i < 3; i++) out[b].fri[i] = real(s.box_friction[b][i]);
    out[b].fri[3] = 0;
  }
  return out;
}
template <typename real> std::vector<real> build_hfield(const QsModel& s) {
  const size_t n = s.terrain_type == QS_TERRAIN_HFIELD ? size_t(s.hf_nrow) * s.hf_ncol : 1;
  std::vector<real> out(n, real(0));
  if (s.terrain_type == QS_TERRAIN_HFIELD) for (size_t i = 0; i < n; i++) out[i] = real(s.hf_data[i]);
  return out;
}

// Features a model / configuration admits (see FEAT_* in qs_env.cuh): a compiled variant may be used iff it asserts a subset.
inline int model_features(const QsModel& m, bool use_imu, int hm_cells) {
  int f = 0;
  f |= m.cone == QS_CONE_PYRAMIDAL ? FEAT_PYR : FEAT_ELL;
  f |= m.terrain_type == QS_TERRAIN_FLAT ? FEAT_FLAT : (m.terrain_type == QS_TERRAIN_HFIELD ? FEAT_HFIELD : FEAT_BOXES);
  if (!use_imu) f |= FEAT_NO_IMU;
  if (hm_cells == 0) f |= FEAT_NO_HM;
  bool mesh = false, caps = false, box = false, cyl = false, lim = false;
  for (int g = 0; g < m.ngeom; g++) {
    mesh |= m.geom_type[g] == QS_GEOM_MESH; caps |= m.geom_type[g] == QS_GEOM_CAPSULE;
    box |= m.geom_type[g] == QS_GEOM_BOX; cyl |= m.geom_type[g] == QS_GEOM_CYLINDER;
  }
  for (int j = 0; j < QS_NJNT; j++) lim |= m.jnt_limited[j] != 0;
  if (!mesh) f |= FEAT_NO_MESH;
  if (!caps) f |= FEAT_NO_CAPSULE;
  if (!box) f |= FEAT_NO_BOX;
  if (!cyl) f |= FEAT_NO_CYL;
  if (m.ngeom <= 32) f |= FEAT_NGEOM32;
  if (!lim) f |= FEAT_NO_LIMITS;
  return f;
}

inline int model_max_dim(const QsModel& s) {
  int md = 1;
  for (int g = 0; g < s.ngeom; g++) {
    const QsGeomParams& gp = s.geom_par[g];
    const int dim = gp.priority == s.floor_par.priority ? std::max(gp.condim, s.floor_par.condim) : (gp.priority > s.floor_par.priority ? gp.condim : s.floor_par.condim);
    md = std::max(md, dim);
  }
  return md;
}

}  // namespace qs
