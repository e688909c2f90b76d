// qs_env.cuh -- per-environment physics of the fused step kernel, one environment per warp.
//
// Everything the reference obtains from `mujoco.mj_step` (gym_quadruped/quadruped_env.py:271) -- forward kinematics,
// composite-rigid-body mass matrix, RNE bias forces, collision against the terrain, soft-constraint build, Newton
// solve, implicit-damping Euler integration -- plus the env-side observation / termination pack
// (quadruped_env.py:1146-1257) is expressed here as warp-cooperative phases:
//   lanes <-> bodies (13), dofs (18), geoms, constraint units; cross-lane traffic through a per-warp shared-memory
//   workspace (`WS`) and warp shuffles; `syncwarp()` between phases.
// The tree sparsity (6-dof base + four independent 3-dof legs) is exploited throughout: M and the Newton Hessian are
// stored as {base 6x6, four 3x6 couplings, four 3x3 leg blocks} and factored by a Schur complement on the leg blocks.
//
// Third-party attribution: comments marked [MJ] name the routine of the MuJoCo physics engine (Google DeepMind, Apache License 2.0;
// the engine behind the reference's `mujoco` dependency, not part of the reference checkout) whose published algorithm a phase
// restates.  No MuJoCo source was copied.
//
// This header is compiled by nvcc for sm_100a (qstep.cu) and, unchanged, by g++ against a 32-lane warp emulator (lock-step fibers)
// (tests/emu) so the exact kernel source can be checked against the fp64 oracle on a CPU-only host.
#pragma once
#include "qs_math.cuh"

namespace qs {

constexpr int NB = 14, NJ = 12, NQ = 19, NV = 18, NU = 12, MAXGEOM = 48;
constexpr int NOBS_BASE = 227;
constexpr int NFL = 12;   // friction-loss units (one per hinge)
constexpr int NLIM = 12;  // joint-limit units (one slot per hinge; lower and upper cannot be active together)

enum { GEOM_PLANE = 0, GEOM_HFIELD = 1, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };

template <typename real> struct alignas(16) Vert4 { real x, y, z, w; };
// static terrain box (random_boxes scene): centre, rotation (row-major), half sizes, bounding radius
template <typename real> struct alignas(16) DBox { real pos[3], mat[9], half[3], rad, fri[4]; };

// Robot + scene constants in the kernel's precision; built on the host from QsModel (qstep.cu: build_dmodel),
// staged into shared memory once per CTA with one TMA bulk copy.
template <typename real> struct alignas(16) DModel {
  real timestep, gravity[3];
  real impratio, tolerance, ls_tolerance, meaninertia;
  real terrain_limits[4];
  real floor_fri[4];
  real imu_pos[4];
  real imu_mat[12];
  real mass_total, robot_radius, hf_lip, pad_r;  // hf_lip: largest slope of any height-field triangle (broad phase of the mesh collider)
  real hf_size[4], hf_pos[4];          // height field: half-x, half-y, z-scale, base; position
  real terr_fri[4], terr_margin, terr_K, terr_B, terr_pad;  // hfield: default parameters; boxes: friction per box (DBox), the rest per scene
  real terr_solimp[8];                         // contact parameters when the scene's boxes out-rank every robot geom (terr_wins)
  real terr_bounds[4];                         // x_max, x_min, y_max, y_min of everything that is not the floor plane, padded by the robot's reach
  real body_pos[NB][3], body_quat[NB][4], body_ipos[NB][3], body_imat[NB][9], body_mass[NB], body_inertia[NB][3], body_iw[NB][2];
  real jnt_pos[NJ][3], jnt_axis[NJ][3], jnt_range[NJ][2], jnt_K[NJ], jnt_B[NJ], jnt_solimp[NJ][5], jnt_margin[NJ];
  real qpos0[20], key_qpos[20];
  real dof_damping[NV], dof_armature[NV], dof_floss[NV], dof_iw[NV], dof_B[NV], dof_R[NV], dof_D[NV];
  real act_clo[NU], act_chi[NU], act_flo[NU], act_fhi[NU];
  real geom_pos[MAXGEOM][3], geom_mat[MAXGEOM][9], geom_size[MAXGEOM][3], geom_bcenter[MAXGEOM][3], geom_bhalf[MAXGEOM][3], geom_rbound[MAXGEOM],
      geom_fri[MAXGEOM][3], geom_margin[MAXGEOM], geom_incmargin[MAXGEOM], geom_K[MAXGEOM], geom_B[MAXGEOM], geom_solimp[MAXGEOM][5];
  int cone, iterations, ls_iterations, ngeom, nvert, terrain_type, nbox, has_imu;
  int hf_nrow, hf_ncol, terr_dim, terr_wins;
  int jnt_limited[NJ];
  int geom_type[MAXGEOM], geom_body[MAXGEOM], geom_leg[MAXGEOM], geom_vertadr[MAXGEOM], geom_vertnum[MAXGEOM], geom_dim[MAXGEOM],
      geom_prio[MAXGEOM];
  int foot_geom[4];
  int pad_i[4];
};

// Per-warp workspace (shared memory).  Arrays with disjoint lifetimes share storage so that 28 warps (= 28 envs) fit in
// one SM's 227 KB next to the staged model:  4096 envs -> 147 CTAs of 28 warps -> a single wave on 148 SMs.
//   kin  (body poses)          : kinematics .. collision / cdof        | hes (Hessian blocks + factors): solver .. Euler
//   tmp  (velocity-stage data) : com_inertia .. mass matrix            | Jc  (contact Jacobians)       : constraints .. solver
//                                                                      | obs (staged observation row)  : after Euler
template <typename real, int NCON, int MAXDIM> struct alignas(16) WS {
  static constexpr int NW = MAXDIM * (MAXDIM + 1) / 2;  // packed symmetric contact weight
  // staged terrain boxes (collision stage): they overlay everything behind tmp.cinert, which is written later in the step
  static constexpr int kStgPad = NV * 6 + NB * 10;
  static constexpr int kUnionWords = (NCON * MAXDIM * 9 > NV * 6 + NB * 22) ? NCON * MAXDIM * 9 : NV * 6 + NB * 22;
  static constexpr int KST = int((kUnionWords - kStgPad) * sizeof(real) / sizeof(DBox<real>));
  // state (internal frame: base xy relative to the per-env origin `org`)
  real qpos[20], qvel[NV], ctrl[NU], warm[NV], applied[6];
  real mu_floor, mu_feet;
  double org[2];
  double tmpd[2];
  // persistent kinematic results
  real com[4], cdof[NV][6], footpos[4][3];
  union {
    struct { real xpos[NB][3], xmat[NB][9], xaxis[NJ][3]; } kin;
    struct { real Hbb[6][6], Hlb[4][3][6], Hll[4][3][3], Ci[4][3][3], Y[4][3][6], SL[6][6], SLinv[8]; } hes;
  };
  union {
    struct { real cdofdot[NV][6], cinert[NB][10], cvel[NB][6], cfrc[NB][6]; } tmp;
    real Jc[NCON][MAXDIM][9];
    real obs[NOBS_BASE + 5];
    struct { real pad_[kStgPad]; DBox<real> box[KST]; } stg;
  };
  unsigned char near_id[32];  // global index of the k-th static box within reach of the robot (k < nnear)
  // block mass matrix: base-base, leg-base, leg-leg
  real Mbb[6][6], Mlb[4][3][6], Mll[4][3][3];
  // dof vectors
  real fsm[NV], asmooth[NV], qacc[NV], fcon[NV], grad[NV], search[NV], Mv[NV], Ma[NV];
  // constraint units: [0,12) friction loss, [12,24) joint limits, then contacts
  real u_ar[NFL + NLIM], u_D[NFL + NLIM], u_sign[NFL + NLIM], u_r[NFL + NLIM], u_F[NFL + NLIM];
  union { real u_v[NFL + NLIM]; real u_W[NFL + NLIM]; };
  int ncon, overflow;
  real c_dist[NCON], c_pos[NCON][3], c_frame[NCON][6], c_fri[NCON][3], c_mu[NCON], c_sign[NCON];
  int c_info[NCON];  // geom | body << 8 | dim << 16
  real c_D[NCON][MAXDIM], c_ar[NCON][MAXDIM], c_r[NCON][MAXDIM], c_F[NCON][MAXDIM];
  union { real c_v[NCON][MAXDIM]; real c_W[NCON][NW]; };
  real sens[8], sens_tmp[24];
};

// Compile-time feature switches of a kernel variant.  A set bit asserts a property of the robot / scene / configuration that the host
// has checked (qstep.cu: model_features), so the code for everything else is not generated at all: the cfg2 step kernel
// (mini_cheetah / flat) carries no elliptic-cone, terrain, capsule / box / cylinder, joint-limit, IMU or ray-cast code.  FEAT = 0 is
// the generic variant that decides everything at run time; every variant produces bit-identical results on a model it admits.
enum : int {
  FEAT_PYR = 1, FEAT_ELL = 2,                       // friction cone known (neither: read from the model)
  FEAT_FLAT = 4, FEAT_HFIELD = 8, FEAT_BOXES = 16,  // scene known: floor plane only / + height field / + static boxes
  FEAT_NO_IMU = 32,                                 // no accelerometer / gyro columns requested: the sensor stage is dropped
  FEAT_NO_MESH = 64, FEAT_NO_CAPSULE = 128, FEAT_NO_BOX = 256, FEAT_NO_CYL = 512,  // robot geom types that do not occur
  FEAT_NGEOM32 = 1024,                              // at most 32 robot collision geoms: a single lane round
  FEAT_NO_LIMITS = 2048,                            // no limited joint
  FEAT_NO_HM = 4096,                                // no fused height-map columns
};
// the specialised step kernels that are compiled (qs_variants.h): one per BASELINE configuration
constexpr int FEAT_CFG2 = FEAT_PYR | FEAT_FLAT | FEAT_NO_IMU | FEAT_NO_HM | FEAT_NO_CAPSULE | FEAT_NO_BOX | FEAT_NO_CYL | FEAT_NGEOM32 | FEAT_NO_LIMITS;  // mini_cheetah / flat
constexpr int FEAT_CFG3 = FEAT_PYR | FEAT_HFIELD | FEAT_NO_IMU | FEAT_NO_MESH | FEAT_NO_CYL | FEAT_NGEOM32;                                              // aliengo / perlin (+ height map)
constexpr int FEAT_CFG4 = FEAT_ELL | FEAT_BOXES | FEAT_NO_IMU | FEAT_NO_HM | FEAT_NO_MESH | FEAT_NO_CYL | FEAT_NGEOM32;                                 // go2 / random_boxes
constexpr int FEAT_CFG5 = FEAT_ELL | FEAT_FLAT | FEAT_NO_HM | FEAT_NO_CAPSULE | FEAT_NO_BOX | FEAT_NO_CYL | FEAT_NGEOM32;                               // hyqreal1 / flat (+ IMU)
// further flat-floor variants for the robots outside the BASELINE configurations
constexpr int FEAT_PYR_FLAT_PRIM = FEAT_PYR | FEAT_FLAT | FEAT_NO_IMU | FEAT_NO_HM | FEAT_NO_MESH | FEAT_NGEOM32;                                       // aliengo, hyqreal2, b2 / flat
constexpr int FEAT_ELL_FLAT_PRIM = FEAT_ELL | FEAT_FLAT | FEAT_NO_IMU | FEAT_NO_HM | FEAT_NO_MESH;                                                      // go2, go1 / flat (condim 6)
constexpr int FEAT_ELL_FLAT_MESH6 = FEAT_ELL | FEAT_FLAT | FEAT_NO_IMU | FEAT_NO_HM | FEAT_NO_CAPSULE | FEAT_NO_BOX | FEAT_NO_CYL | FEAT_NGEOM32;      // spot / flat (condim 6)

template <typename real, int NCON, int MAXDIM, int FEAT = 0> struct Env {
  using N = Num<real>;
  using W = WS<real, NCON, MAXDIM>;
  static constexpr int NW = W::NW;
  static constexpr bool kLimits = !(FEAT & FEAT_NO_LIMITS);
  static constexpr int NSC = kLimits ? NFL + NLIM : NFL;  // scalar constraint units in use
  const DModel<real>& m;
  W& w;
  const Vert4<real>* vert;
  const real* hf;             // height-field samples [nrow][ncol] in [0,1] (global memory), or nullptr
  const DBox<real>* boxes;    // static terrain boxes (global memory), or nullptr
  const int lane;
  int solver_iter, ls_evals;
  bool solver_maxed;
  float* bias_out;  // optional destination for qfrc_bias (accessor dump), else nullptr
#ifdef QS_PROF
  unsigned tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // diagnostic builds: cycles per solver sub-phase
  long long tlast = 0;
#define QS_T0() tlast = clock64()
#define QS_TACC(k) do { const long long t_ = clock64(); tacc[k] += unsigned(t_ - tlast); tlast = t_; } while (0)
#else
#define QS_T0() do { } while (0)
#define QS_TACC(k) do { } while (0)
#endif

  int tri;  // (i0, j0, i1, j1), 4 bits each: rows / columns of entries e = lane and e = lane + 32 of a row-major lower triangle

  QS_DEV Env(const DModel<real>& m_, W& w_, const Vert4<real>* v_, int lane_)
      : m(m_), w(w_), vert(v_), hf(nullptr), boxes(nullptr), lane(lane_), solver_iter(0), ls_evals(0), solver_maxed(false), bias_out(nullptr) {
    // per lane: i0 | j0<<4 | i1<<8 | j1<<12 | dof_leg<<16 | dof_k<<18 | (lane/6)<<20 | (lane%6)<<23, with (i, j) the row / column of
    // entry e = lane (and e = lane + 32) of a row-major lower triangle -- tabulated (checked by tests/test_abi_and_host.py)
    static constexpr int kLaneRoles[32] = {18176, 8410881, 16803601, 25196290, 33556498, 41949218, 1058819, 9713683, 18368547, 26302515, 34957316, 43612180, 2263076, 10881332, 19536196, 27470085, 36124949, 44779813, 3168565, 11561285, 19954005, 28346630, 36739350, 45091366, 4201014, 12593734, 20986454, 29379174, 37771783, 46164503, 5274151, 13666871};
    tri = kLaneRoles[lane_ & 31];
  }
  // leg / joint-in-leg of dof lane 6..17 (0 elsewhere); lane / 6 and lane % 6
  QS_DEV int dof_leg() const { return (tri >> 16) & 3; }
  QS_DEV int dof_k() const { return (tri >> 18) & 3; }
  QS_DEV static int tri_dof_leg(int t) { return (t >> 16) & 3; }
  QS_DEV static int tri_dof_k(int t) { return (t >> 18) & 3; }

  QS_DEV bool cone_is_pyramidal() const { return (FEAT & FEAT_PYR) ? true : ((FEAT & FEAT_ELL) ? false : m.cone == 0); }
  QS_DEV int ttype() const { return (FEAT & FEAT_FLAT) ? 0 : ((FEAT & FEAT_HFIELD) ? 1 : ((FEAT & FEAT_BOXES) ? 2 : m.terrain_type)); }
  QS_DEV bool has_imu() const { return (FEAT & FEAT_NO_IMU) ? false : m.has_imu != 0; }
  // robot geom type t can occur in this variant
  QS_DEV static constexpr bool type_on(int t) {
    return t == GEOM_MESH ? !(FEAT & FEAT_NO_MESH) : (t == GEOM_CAPSULE ? !(FEAT & FEAT_NO_CAPSULE) : (t == GEOM_BOX ? !(FEAT & FEAT_NO_BOX)
         : (t == GEOM_CYLINDER ? !(FEAT & FEAT_NO_CYL) : true)));
  }
  unsigned cm_acc = 0, im_acc = 0;  // per-lane share of the contact / invalid-contact body masks, collected when a contact is DETECTED
                                    // (before the NCON cap), so that termination stays exact even if the contact buffer overflows
  real pen_acc = 0;                 // per-lane share of max |dist| over the calf-body contacts, collected the same way: the reset lift
                                    // loop raises the robot by 1.1 x this (quadruped_env.py:381-383), also when contacts were dropped
  QS_DEV void note_contact(int g, real dist) {
    const int b = m.geom_body[g];
    if (b >= 2 && (b - 2) % 3 == 2) { cm_acc |= 1u << ((b - 2) / 3); pen_acc = N::max(pen_acc, N::abs(dist)); }
    else im_acc |= 1u << b;
  }
  bool terrain_on = true;  // false: the base is out of reach of everything but the floor plane (internal frame re-centred like 'flat')
  bool calf_only = false;  // collision stage restricted to the calf-body geoms (the reset lift loop looks at nothing else)
  QS_DEV bool geom_on(int g) const {
    if (g >= m.ngeom) return false;
    const int b = m.geom_body[g];
    return !calf_only || (b >= 2 && (b - 2) % 3 == 2);
  }

  QS_DEV static int info_geom(int info) { return info & 0xff; }
  QS_DEV static int info_body(int info) { return (info >> 8) & 0xff; }
  QS_DEV static int info_dim(int info) { return (info >> 16) & 0xff; }
  QS_DEV static int info_leg(int info) { return (info >> 24) & 7; }  // leg of the contact body, 7: base / world side only
  QS_DEV static bool info_terrain_wins(int info) { return (info >> 27) & 1; }  // parameters of a higher-priority terrain box apply
  QS_DEV static int widx(int a, int b) { return a <= b ? a * MAXDIM - (a * (a - 1)) / 2 + (b - a) : b * MAXDIM - (b * (b - 1)) / 2 + (a - b); }

  // ------------------------------------------------------------------ position stage
  // [MJ] mj_kinematics: lane j<12 walks the chain base -> ... -> body j+2 on its own, lane 12 owns the base.
  // Joint anchors coincide with the body origins in every supported model (jnt_pos = 0, checked at model build).
  QS_DEV void kinematics() {
    real qb[4] = {w.qpos[3], w.qpos[4], w.qpos[5], w.qpos[6]};
    quat_normalize(qb);
    real Rb[9];
    quat_to_mat(Rb, qb);
    if (lane == 12) {
      for (int i = 0; i < 3; i++) w.kin.xpos[1][i] = w.qpos[i];
      for (int i = 0; i < 9; i++) w.kin.xmat[1][i] = Rb[i];
    } else if (lane == 13) {
      for (int i = 0; i < 3; i++) w.kin.xpos[0][i] = 0;
      for (int i = 0; i < 9; i++) w.kin.xmat[0][i] = (i % 4 == 0) ? real(1) : real(0);
    } else if (lane < 12) {
      const int l = lane / 3, k = lane % 3;
      real pos[3] = {w.qpos[0], w.qpos[1], w.qpos[2]}, quat[4] = {qb[0], qb[1], qb[2], qb[3]}, R[9], axis[3] = {0, 0, 0};
      for (int i = 0; i < 9; i++) R[i] = Rb[i];
      for (int t = 0; t < 3; t++) {
        if (t > k) break;
        const int b = 2 + 3 * l + t, j = b - 2;
        real tmp[3], q2[4];
        mul_mv(tmp, R, m.body_pos[b]);
        for (int i = 0; i < 3; i++) pos[i] += tmp[i];
        quat_mul(q2, quat, m.body_quat[b]);
        if (t == k) { quat_to_mat(R, q2); mul_mv(axis, R, m.jnt_axis[j]); }
        real s, c;
        N::sincos(real(0.5) * (w.qpos[7 + j] - m.qpos0[7 + j]), &s, &c);
        real qloc[4] = {c, s * m.jnt_axis[j][0], s * m.jnt_axis[j][1], s * m.jnt_axis[j][2]};
        quat_mul(quat, q2, qloc);
        quat_normalize(quat);
        quat_to_mat(R, quat);
      }
      const int b = lane + 2;
      for (int i = 0; i < 3; i++) { w.kin.xpos[b][i] = pos[i]; w.kin.xaxis[lane][i] = axis[i]; }
      for (int i = 0; i < 9; i++) w.kin.xmat[b][i] = R[i];
    }
    syncwarp();
  }

  // [MJ] mj_comPos: subtree centre of mass and per-body spatial inertias about it; body lanes 0..12 (body = lane+1)
  QS_DEV void com_inertia() {
    const int b = lane + 1;
    const bool isb = lane < 13;
    const int bb = isb ? b : 1;
    real xipos[3], Ri[9], tmp[3];
    mul_mv(tmp, w.kin.xmat[bb], m.body_ipos[bb]);
    for (int i = 0; i < 3; i++) xipos[i] = w.kin.xpos[bb][i] + tmp[i];
    {
      const real* A = w.kin.xmat[bb];
      const real* Bm = m.body_imat[bb];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ri[3 * r + c] = A[3 * r] * Bm[c] + A[3 * r + 1] * Bm[3 + c] + A[3 * r + 2] * Bm[6 + c];
    }
    const real mass = isb ? m.body_mass[bb] : real(0);
    real c[3];
    const real inv_mass = N::rcp(m.mass_total);
    for (int i = 0; i < 3; i++) c[i] = warp_sum(mass * xipos[i]) * inv_mass;
    if (isb) {
      const real* in = m.body_inertia[bb];
      real* I = w.tmp.cinert[b];
      real o[3] = {xipos[0] - c[0], xipos[1] - c[1], xipos[2] - c[2]};
      real A00 = Ri[0] * in[0] * Ri[0] + Ri[1] * in[1] * Ri[1] + Ri[2] * in[2] * Ri[2];
      real A11 = Ri[3] * in[0] * Ri[3] + Ri[4] * in[1] * Ri[4] + Ri[5] * in[2] * Ri[5];
      real A22 = Ri[6] * in[0] * Ri[6] + Ri[7] * in[1] * Ri[7] + Ri[8] * in[2] * Ri[8];
      real A01 = Ri[0] * in[0] * Ri[3] + Ri[1] * in[1] * Ri[4] + Ri[2] * in[2] * Ri[5];
      real A02 = Ri[0] * in[0] * Ri[6] + Ri[1] * in[1] * Ri[7] + Ri[2] * in[2] * Ri[8];
      real A12 = Ri[3] * in[0] * Ri[6] + Ri[4] * in[1] * Ri[7] + Ri[5] * in[2] * Ri[8];
      real oo = dot3(o, o);
      I[0] = A00 + mass * (oo - o[0] * o[0]); I[1] = A11 + mass * (oo - o[1] * o[1]); I[2] = A22 + mass * (oo - o[2] * o[2]);
      I[3] = A01 - mass * o[0] * o[1]; I[4] = A02 - mass * o[0] * o[2]; I[5] = A12 - mass * o[1] * o[2];
      I[6] = mass * o[0]; I[7] = mass * o[1]; I[8] = mass * o[2]; I[9] = mass;
    }
    if (lane == 13) for (int i = 0; i < 3; i++) w.com[i] = c[i];
    syncwarp();
  }

  // [MJ] cdof (mj_comPos); dof lanes 0..17
  QS_DEV void cdof() {
    if (lane < NV) {
      const int d = lane;
      real ax[3], off[3], cd[6];
      if (d < 3) {
        cd[0] = cd[1] = cd[2] = 0; cd[3] = d == 0; cd[4] = d == 1; cd[5] = d == 2;
      } else {
        if (d < 6) {
          const int kk = d - 3;
          ax[0] = w.kin.xmat[1][kk]; ax[1] = w.kin.xmat[1][3 + kk]; ax[2] = w.kin.xmat[1][6 + kk];
          for (int i = 0; i < 3; i++) off[i] = w.com[i] - w.kin.xpos[1][i];
        } else {
          for (int i = 0; i < 3; i++) { ax[i] = w.kin.xaxis[d - 6][i]; off[i] = w.com[i] - w.kin.xpos[d - 4][i]; }
        }
        cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
        cross3(cd + 3, ax, off);
      }
      for (int i = 0; i < 6; i++) w.cdof[d][i] = cd[i];
    }
    syncwarp();
  }

  // [MJ] mj_crb: composite inertias (in place, by shuffles) then M in block form; must run after bias_and_smooth
  QS_DEV void mass_matrix() {
    {
      const bool isb = lane < 13;
      const int b = lane + 1;
      const int k = (lane >= 1 && isb) ? (lane - 1) % 3 : 3;
      real cr[10];
      const real* ci = w.tmp.cinert[isb ? b : 1];
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const real v = isb ? ci[i] : real(0);
        const real tot = warp_sum(v);
        const real t1 = shfl(v, (lane + 1) & 31), t2 = shfl(v, (lane + 2) & 31);
        cr[i] = (lane == 0) ? tot : v + (k <= 1 ? t1 : real(0)) + (k == 0 ? t2 : real(0));
      }
      syncwarp();
      if (isb) {
        real* co = w.tmp.cinert[b];
#pragma unroll
        for (int i = 0; i < 10; i++) co[i] = cr[i];
      }
    }
    syncwarp();
    if (lane < NV) {
      const int d = lane, body = d < 6 ? 1 : d - 4;
      real buf[6], cd[6];
      for (int i = 0; i < 6; i++) cd[i] = w.cdof[d][i];
      mul_inert_vec(buf, w.tmp.cinert[body], cd);
      if (d < 6) {
        for (int j = 0; j <= d; j++) {
          real v = 0;
          for (int i = 0; i < 6; i++) v += w.cdof[j][i] * buf[i];
          if (j == d) v += m.dof_armature[d];
          w.Mbb[d][j] = v; w.Mbb[j][d] = v;
        }
      } else {
        const int l = (d - 6) / 3, k = (d - 6) % 3;
        for (int j = 0; j < 6; j++) {
          real v = 0;
          for (int i = 0; i < 6; i++) v += w.cdof[j][i] * buf[i];
          w.Mlb[l][k][j] = v;
        }
        for (int k2 = 0; k2 <= k; k2++) {
          real v = 0;
          for (int i = 0; i < 6; i++) v += w.cdof[6 + 3 * l + k2][i] * buf[i];
          if (k2 == k) v += m.dof_armature[d];
          w.Mll[l][k][k2] = v; w.Mll[l][k2][k] = v;
        }
      }
    }
    syncwarp();
  }

  // ------------------------------------------------------------------ velocity / force stage
  // [MJ] mj_comVel + mj_rne + passive + actuation; body lanes then dof lanes
  QS_DEV void bias_and_smooth() {
    const int b = lane + 1;
    const bool isb = lane < 13;
    const int l = (lane >= 1 && isb) ? (lane - 1) / 3 : 0, k = (lane >= 1 && isb) ? (lane - 1) % 3 : -1;
    real cv[6] = {0, 0, 0, 0, 0, 0};
    if (isb) {
      // base: translations, then rotations (cdofdot of the rotations uses the velocity after the translations)
      real vt[6] = {0, 0, 0, w.qvel[0], w.qvel[1], w.qvel[2]};
      for (int i = 0; i < 6; i++) cv[i] = vt[i];
      for (int d = 3; d < 6; d++) for (int i = 0; i < 6; i++) cv[i] += w.cdof[d][i] * w.qvel[d];
      if (lane == 0) {
        for (int d = 0; d < 3; d++) for (int i = 0; i < 6; i++) w.tmp.cdofdot[d][i] = 0;
        for (int d = 3; d < 6; d++) { real t[6]; cross_motion(t, vt, w.cdof[d]); for (int i = 0; i < 6; i++) w.tmp.cdofdot[d][i] = t[i]; }
      } else {
        for (int t = 0; t <= k; t++) {
          const int d = 6 + 3 * l + t;
          if (t == k) { real cdd[6]; cross_motion(cdd, cv, w.cdof[d]); for (int i = 0; i < 6; i++) w.tmp.cdofdot[d][i] = cdd[i]; }
          for (int i = 0; i < 6; i++) cv[i] += w.cdof[d][i] * w.qvel[d];
        }
      }
      for (int i = 0; i < 6; i++) w.tmp.cvel[b][i] = cv[i];
    }
    syncwarp();
    real f[6] = {0, 0, 0, 0, 0, 0};
    if (isb) {
      real ca[6] = {0, 0, 0, -m.gravity[0], -m.gravity[1], -m.gravity[2]};
      for (int d = 3; d < 6; d++) for (int i = 0; i < 6; i++) ca[i] += w.tmp.cdofdot[d][i] * w.qvel[d];
      for (int t = 0; t <= k; t++) { const int d = 6 + 3 * l + t; for (int i = 0; i < 6; i++) ca[i] += w.tmp.cdofdot[d][i] * w.qvel[d]; }
      if (lane == 0 && has_imu()) {
        // IMU pre-computation (everything except the cdof*qacc term), kept because kin/tmp storage is recycled by the solver
        real* st = w.sens_tmp;
        for (int i = 0; i < 6; i++) { st[i] = cv[i]; st[6 + i] = ca[i]; }
        real tmp[3];
        mul_mv(tmp, w.kin.xmat[1], m.imu_pos);
        for (int i = 0; i < 3; i++) st[12 + i] = w.kin.xpos[1][i] + tmp[i] - w.com[i];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) st[15 + 3 * r + c] = w.kin.xmat[1][3 * r] * m.imu_mat[c] + w.kin.xmat[1][3 * r + 1] * m.imu_mat[3 + c] + w.kin.xmat[1][3 * r + 2] * m.imu_mat[6 + c];
      }
      real I[10], t1[6], t2[6];
      for (int i = 0; i < 10; i++) I[i] = w.tmp.cinert[b][i];
      mul_inert_vec(t1, I, ca);
      mul_inert_vec(t2, I, cv);
      cross_force(f, cv, t2);
      for (int i = 0; i < 6; i++) f[i] += t1[i];
    }
    const int kk = (lane >= 1 && isb) ? k : 3;
    real s6[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      real tot = warp_sum(f[i]);
      real t1 = shfl(f[i], (lane + 1) & 31), t2 = shfl(f[i], (lane + 2) & 31);
      s6[i] = (lane == 0) ? tot : f[i] + (kk <= 1 ? t1 : real(0)) + (kk == 0 ? t2 : real(0));
    }
    if (isb) {
      real* cf = w.tmp.cfrc[b];
#pragma unroll
      for (int i = 0; i < 6; i++) cf[i] = s6[i];
    }
    syncwarp();
    if (lane < NV) {
      const int d = lane, body = d < 6 ? 1 : d - 4;
      real s = 0;
      for (int i = 0; i < 6; i++) s += w.cdof[d][i] * w.tmp.cfrc[body][i];
      if (bias_out) bias_out[d] = float(s);
      real act;
      if (d < 6) act = w.applied[d];
      else {
        const int a = d - 6;
        real c = w.ctrl[a];
        c = N::min(N::max(c, m.act_clo[a]), m.act_chi[a]);
        c = N::min(N::max(c, m.act_flo[a]), m.act_fhi[a]);
        act = c;
      }
      w.fsm[d] = -m.dof_damping[d] * w.qvel[d] - s + act;
    }
    syncwarp();
  }

  // Accessor tables for mj_jac (rotational part) and mj_jacDot users (quadruped_env.py:681-797): per foot l and dof d, the
  // rotational Jacobian and the time derivatives of both Jacobians of the calf body at the foot point, as [4][3][18] floats each.
  // [MJ] mj_jacDot: free-joint rotations rebuild cdof_dot from the full body velocity.  Valid right after bias_and_smooth().
  QS_DEV void dump_jacobian_tables(float* jacr, float* jacp_dot, float* jacr_dot) const {
    for (int it = lane; it < 4 * NV; it += 32) {
      const int l = it / NV, d = it % NV, body = 4 + 3 * l;
      real jr[3] = {0, 0, 0}, jpd[3] = {0, 0, 0}, jrd[3] = {0, 0, 0};
      if (d < 6 || (d - 6) / 3 == l) {
        const real off[3] = {w.footpos[l][0] - w.com[0], w.footpos[l][1] - w.com[1], w.footpos[l][2] - w.com[2]};
        real cv[6], cd[6], cdd[6] = {0, 0, 0, 0, 0, 0}, pvel[3], t1[3], t2[3];
        for (int i = 0; i < 6; i++) { cv[i] = w.tmp.cvel[body][i]; cd[i] = w.cdof[d][i]; }
        cross3(t1, cv, off);
        for (int i = 0; i < 3; i++) pvel[i] = cv[3 + i] + t1[i];
        if (d >= 6) for (int i = 0; i < 6; i++) cdd[i] = w.tmp.cdofdot[d][i];
        else if (d >= 3) { real cb[6]; for (int i = 0; i < 6; i++) cb[i] = w.tmp.cvel[1][i]; cross_motion(cdd, cb, cd); }
        cross3(t1, cdd, off);
        cross3(t2, cd, pvel);
        for (int i = 0; i < 3; i++) { jr[i] = cd[i]; jpd[i] = cdd[3 + i] + t1[i] + t2[i]; jrd[i] = cdd[i]; }
      }
      for (int i = 0; i < 3; i++) {
        jacr[(l * 3 + i) * NV + d] = float(jr[i]); jacp_dot[(l * 3 + i) * NV + d] = float(jpd[i]); jacr_dot[(l * 3 + i) * NV + d] = float(jrd[i]);
      }
    }
  }

  // ------------------------------------------------------------------ structured linear algebra
  // y = A x for a block matrix (bb, lb, ll); valid on dof lanes (< NV), x read from shared memory
  QS_DEV static real block_matvec(const real (*Abb)[6], const real (*Alb)[3][6], const real (*All)[3][3], const real* x, const int lane, const int tri) {
    // both row shapes are evaluated on clamped indices and the lane keeps its own: a divergent warp would run both anyway
    const bool isbase = lane < 6;
    const int l = tri_dof_leg(tri), k = tri_dof_k(tri), bl = isbase ? lane : 0;
    const real* row = isbase ? Abb[bl] : Alb[l][k];
    real s = 0, sb = 0, sl = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) s += row[j] * x[j];
#pragma unroll
    for (int q = 0; q < 12; q++) sb += (&Alb[0][0][0])[6 * q + bl] * x[6 + q];
#pragma unroll
    for (int k2 = 0; k2 < 3; k2++) sl += All[l][k][k2] * x[6 + 3 * l + k2];
    return s + (isbase ? sb : sl);
  }
  // (M x)[lane] with the mass blocks of the workspace: one out-of-line copy for the five call sites of a step
  QS_NOINLINE static real mass_matvec(const W& w, const real* x, const int lane, const int tri) {
    return block_matvec(w.Mbb, w.Mlb, w.Mll, x, lane, tri);
  }

  // Factor the block matrix in w.hes (Hbb, Hlb, Hll): leg blocks C_l -> explicit inverses Ci, Y_l = C_l^-1 B_l,
  // Schur complement S = A - sum B_l^T Y_l -> Cholesky SL (lower) with reciprocal diagonal.
  QS_NOINLINE static void factor_H(W& w, const int lane, const int tri) {
    auto& h = w.hes;
    if (lane < 24) {
      const int l = (tri >> 20) & 7, c = (tri >> 23) & 7;
      // LDL^T of the 3x3 leg block, done redundantly by the 6 column lanes of a leg
      const real c00 = h.Hll[l][0][0], c10 = h.Hll[l][1][0], c20 = h.Hll[l][2][0], c11 = h.Hll[l][1][1], c21 = h.Hll[l][2][1], c22 = h.Hll[l][2][2];
      const real id0 = N::rcp(c00), l10 = c10 * id0, l20 = c20 * id0;
      const real d1 = c11 - l10 * c10, id1 = N::rcp(d1);
      const real l21 = (c21 - l20 * c10) * id1;
      const real d2 = c22 - l20 * c20 - l21 * l21 * d1, id2 = N::rcp(d2);
      real b0 = h.Hlb[l][0][c], b1 = h.Hlb[l][1][c], b2 = h.Hlb[l][2][c];
      real z0 = b0, z1 = b1 - l10 * z0, z2 = b2 - l20 * z0 - l21 * z1;
      real y2 = z2 * id2, y1 = z1 * id1 - l21 * y2, y0 = z0 * id0 - l10 * y1 - l20 * y2;
      h.Y[l][0][c] = y0; h.Y[l][1][c] = y1; h.Y[l][2][c] = y2;
      if (c < 3) {  // column c of the explicit inverse
        real e0 = c == 0, e1 = c == 1, e2 = c == 2;
        real u0 = e0, u1 = e1 - l10 * u0, u2 = e2 - l20 * u0 - l21 * u1;
        real v2 = u2 * id2, v1 = u1 * id1 - l21 * v2, v0 = u0 * id0 - l10 * v1 - l20 * v2;
        h.Ci[l][0][c] = v0; h.Ci[l][1][c] = v1; h.Ci[l][2][c] = v2;
      }
    }
    syncwarp();
    // Schur complement entries (lower triangle) spread over 21 lanes, parked in SL
    if (lane < 21) {
      const int i = tri & 15, j = (tri >> 4) & 15;
      real s = h.Hbb[i][j];
      for (int l = 0; l < 4; l++)
        for (int k = 0; k < 3; k++) s -= h.Hlb[l][k][i] * h.Y[l][k][j];
      h.SL[i][j] = s;
    }
    syncwarp();
    // every lane runs the 6x6 Cholesky redundantly in registers (6 dependent pivots, no further synchronisation)
    real S[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) S[i][j] = h.SL[i][j];
    syncwarp();
    real inv[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      real s = S[j][j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= S[j][k] * S[j][k];
      s = N::max(s, N::minval);
      inv[j] = N::rsqrt(s);
      S[j][j] = s * inv[j];
#pragma unroll
      for (int i = j + 1; i < 6; i++) {
        real t = S[i][j];
#pragma unroll
        for (int k = 0; k < j; k++) t -= S[i][k] * S[j][k];
        S[i][j] = t * inv[j];
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) h.SL[i][j] = S[i][j];
        h.SLinv[i] = inv[i];
      }
    }
    syncwarp();
  }

  // solve (factored H) x = rhs for shared-memory vectors rhs, out (may alias); out = scale * x; all lanes participate
  QS_NOINLINE static void solve_H(W& w, const int lane, const int tri, const real* rhs, real* out, const real scale) {
    auto& h = w.hes;
    const int l = tri_dof_leg(tri), k = tri_dof_k(tri);
    const bool isbase = lane < 6, isleg = lane >= 6 && lane < NV;
    // base lanes: t = rhs_b - sum_l Y_l^T rhs_l ; leg lanes: rl = (C_l^-1 rhs_l)_k   (both forms evaluated branch-free on clamped indices)
    const int bl = isbase ? lane : 0;
    real t = rhs[bl], rl = 0;
#pragma unroll
    for (int q = 0; q < 12; q++) t -= (&h.Y[0][0][0])[6 * q + bl] * rhs[6 + q];
#pragma unroll
    for (int k2 = 0; k2 < 3; k2++) rl += h.Ci[l][k][k2] * rhs[6 + 3 * l + k2];
    real tb[6], xb[6];
#pragma unroll
    for (int i = 0; i < 6; i++) tb[i] = shfl(t, i);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      real s = tb[i];
#pragma unroll
      for (int q = 0; q < i; q++) s -= h.SL[i][q] * xb[q];
      xb[i] = s * h.SLinv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      real s = xb[i];
#pragma unroll
      for (int q = i + 1; q < 6; q++) s -= h.SL[q][i] * xb[q];
      xb[i] = s * h.SLinv[i];
    }
    // leg lanes: x_l = C^-1 rhs_l - Y_l x_b ; base lanes pick their own component without indexing the register array
    real sl = rl, sb = xb[0];
#pragma unroll
    for (int j = 0; j < 6; j++) { sl -= h.Y[l][k][j] * xb[j]; if (j > 0) sb = (lane == j) ? xb[j] : sb; }
    syncwarp();  // rhs fully consumed before out (possibly the same array) is written
    if (lane < NV) out[lane] = scale * (isbase ? sb : sl);
    syncwarp();
  }

  // H = M + h_damp * diag(damping).  The hes blocks (Hbb, Hlb, Hll) and the mass blocks (Mbb, Mlb, Mll) are laid out identically
  // (144 contiguous words each), so this is a flat copy followed by 18 diagonal updates.
  QS_DEV void copy_M_to_H(real h_damp) {
    real* H = &w.hes.Hbb[0][0];
    const real* M = &w.Mbb[0][0];
#pragma unroll
    for (int r = 0; r < 5; r++) { const int e = lane + 32 * r; if (r < 4 || lane < 16) H[e] = M[e]; }
    if (h_damp != 0) {
      syncwarp();
      if (lane < 6) w.hes.Hbb[lane][lane] += h_damp * m.dof_damping[lane];
      else if (lane < NV) { const int l = dof_leg(), k = dof_k(); w.hes.Hll[l][k][k] += h_damp * m.dof_damping[lane]; }
    }
    syncwarp();
  }

  // ------------------------------------------------------------------ collision with the terrain
  // wg: world geom of the contact (WG_FLOOR, WG_HFIELD, or the index of a static box)
  enum { WG_FLOOR = -1, WG_HFIELD = -2 };
  QS_DEV static void contact_friction(const W& w, const DModel<real>& m, const DBox<real>* boxes, int g, int wg, real* fri) {
    const bool world_is_floor = wg == WG_FLOOR;
    real gf[3] = {m.geom_fri[g][0], m.geom_fri[g][1], m.geom_fri[g][2]};
    if (m.geom_leg[g] >= 0 && w.mu_feet >= 0) { gf[0] = w.mu_feet; gf[1] = real(0.005); gf[2] = 0; }   // quadruped_env.py:1290-1296
    real wf[3] = {world_is_floor ? m.floor_fri[0] : m.terr_fri[0], world_is_floor ? m.floor_fri[1] : m.terr_fri[1], world_is_floor ? m.floor_fri[2] : m.terr_fri[2]};
    if (wg >= 0) for (int i = 0; i < 3; i++) wf[i] = boxes[wg].fri[i];
    if (world_is_floor && w.mu_floor >= 0) { wf[0] = w.mu_floor; wf[1] = real(0.005); wf[2] = 0; }
    const int prio = (wg >= 0 && m.terr_wins) ? -1 : m.geom_prio[g];
    for (int i = 0; i < 3; i++) {
      real f = prio == 0 ? N::max(gf[i], wf[i]) : (prio > 0 ? gf[i] : wf[i]);
      fri[i] = N::max(f, real(1e-5));  // [MJ] mjMINMU
    }
  }

  // store one contact (called by a single lane); normal given, yhint optional second-axis guess.  One shared out-of-line body
  // (scalar arguments, no `this`): the three collider stages that call it would otherwise each inline ~1.1 k instructions that
  // run a handful of times per env-step.
  QS_DEV void store_contact(int slot, int g, real sign, real dist, const real* pos, const real* normal, const real* yhint, int wg) {
    store_contact_impl(w, m, boxes, slot, g, sign, dist, pos[0], pos[1], pos[2], normal[0], normal[1], normal[2], yhint ? yhint[0] : real(0),
                       yhint ? yhint[1] : real(0), yhint ? yhint[2] : real(0), wg);
  }
  QS_NOINLINE static void store_contact_impl(W& w, const DModel<real>& m, const DBox<real>* boxes, int slot, int g, real sign, real dist, real px,
                                             real py, real pz, real nx, real ny, real nz, real yx, real yy, real yz, int wg) {
    w.c_dist[slot] = dist; w.c_sign[slot] = sign;
    const bool tw = wg >= 0 && m.terr_wins;
    const int gb = m.geom_body[g], gleg = gb >= 2 ? (gb - 2) / 3 : 7;
    w.c_info[slot] = g | (gb << 8) | ((tw ? m.terr_dim : m.geom_dim[g]) << 16) | (gleg << 24) | (tw ? (1 << 27) : 0);
    real f[6] = {nx, ny, nz, yx, yy, yz};
    w.c_pos[slot][0] = px; w.c_pos[slot][1] = py; w.c_pos[slot][2] = pz;
    // [MJ] mju_makeFrame (third axis = normal x second axis, rebuilt on demand)
    if (dot3(f + 3, f + 3) < real(0.25)) { f[3] = f[4] = f[5] = 0; if (f[1] < real(0.5) && f[1] > real(-0.5)) f[4] = 1; else f[5] = 1; }
    real dd = dot3(f, f + 3);
    for (int i = 0; i < 3; i++) f[3 + i] -= dd * f[i];
    real inv = N::rsqrt(dot3(f + 3, f + 3));
    for (int i = 0; i < 3; i++) f[3 + i] *= inv;
    for (int i = 0; i < 6; i++) w.c_frame[slot][i] = f[i];
    real fri[3];
    contact_friction(w, m, boxes, g, wg, fri);
    for (int i = 0; i < 3; i++) w.c_fri[slot][i] = fri[i];
  }
  // row k (0 normal, 1, 2 tangents) of the contact frame
  QS_DEV void frame_row(int c, int k, real* out) const {
    const real* f = w.c_frame[c];
    if (k == 0) { out[0] = f[0]; out[1] = f[1]; out[2] = f[2]; }
    else if (k == 1) { out[0] = f[3]; out[1] = f[4]; out[2] = f[5]; }
    else cross3(out, f, f + 3);
  }

  // ------------------------------------------------------------------ terrain beyond the floor plane
  // height of the perlin field under (x, y) and the unit normal of the triangle there; cells split along (0,0)-(1,1) [MJ]
  QS_DEV bool hfield_height(real x, real y, real& z, real* n) const {
    const real sx = m.hf_size[0], sy = m.hf_size[1], sz = m.hf_size[2];
    const real lx = x - m.hf_pos[0], ly = y - m.hf_pos[1];
    if (lx < -sx || lx > sx || ly < -sy || ly > sy) return false;
    const int nc = m.hf_ncol, nr = m.hf_nrow;
    const real dx = 2 * sx / real(nc - 1), dy = 2 * sy / real(nr - 1);
    int c = int(N::floor((lx + sx) / dx)), r = int(N::floor((ly + sy) / dy));
    c = c > nc - 2 ? nc - 2 : (c < 0 ? 0 : c);
    r = r > nr - 2 ? nr - 2 : (r < 0 ? 0 : r);
    const real u = (lx + sx - c * dx) / dx, v = (ly + sy - r * dy) / dy;
    const real z00 = sz * hf[r * nc + c], z10 = sz * hf[r * nc + c + 1], z01 = sz * hf[(r + 1) * nc + c], z11 = sz * hf[(r + 1) * nc + c + 1];
    real gx, gy;
    if (u >= v) { z = z00 + u * (z10 - z00) + v * (z11 - z10); gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; }
    else { z = z00 + u * (z11 - z01) + v * (z01 - z00); gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; }
    z += m.hf_pos[2];
    const real inv = N::rsqrt(gx * gx + gy * gy + 1);
    n[0] = -gx * inv; n[1] = -gy * inv; n[2] = inv;
    return true;
  }

  // candidate terrain contact kept in registers by the geom's lane (the 4 deepest per geom survive)
  struct Cand { real dist, pos[3], nrm[3], sign; int wg; };
  QS_DEV static void cand_insert(Cand* list, int& n, const Cand& c) {
    int k = n < 4 ? n : 4;
    if (n >= 4 && !(c.dist < list[3].dist)) return;
    if (k == 4) k = 3;
    while (k > 0 && c.dist < list[k - 1].dist) { list[k] = list[k - 1]; k--; }
    list[k] = c;
    if (n < 4) n++;
  }
  // point feature (centre p, radius r) against the height field or the candidate boxes; robot_first: sphere / capsule sort
  // before box in the engine's geom ordering, so the robot geom is geom1 and the normal points from it into the box
  QS_DEV void point_vs_terrain(const real* p, real r, real margin, bool robot_first, unsigned boxmask, Cand* list, int& n) const {
    if (ttype() == 1) {
      real z, nn[3];
      if (!hfield_height(p[0], p[1], z, nn)) return;
      const real dist = (p[2] - z) * nn[2] - r;
      if (dist > margin) return;
      Cand c;
      c.dist = dist; c.sign = 1; c.wg = WG_HFIELD;
      for (int i = 0; i < 3; i++) { c.nrm[i] = nn[i]; c.pos[i] = p[i] - nn[i] * (r + real(0.5) * dist); }
      cand_insert(list, n, c);
    } else if (ttype() == 2) {
      unsigned mask = boxmask;
      while (mask) {
        const int k = ctz(mask);
        mask &= mask - 1;
        point_vs_box(p, r, margin, robot_first, k, list, n);
      }
    }
  }
  // point feature (centre p, radius r) against the k-th near box
  QS_DEV void point_vs_box(const real* p, real r, real margin, bool robot_first, int k, Cand* list, int& n) const {
    {
      {
        {
          const DBox<real>& bx = near_box(k);
          const int b = w.near_id[k];
          const real rel[3] = {p[0] - bx.pos[0], p[1] - bx.pos[1], p[2] - bx.pos[2]};
          const real reach = bx.rad + r + real(0.01);
          if (dot3(rel, rel) > reach * reach) return;
          real q[3], cl, dl[3], nl[3] = {0, 0, 0}, dist;
          mul_mtv(q, bx.mat, rel);
          bool inside = true;
          for (int i = 0; i < 3; i++) { cl = q[i] < -bx.half[i] ? -bx.half[i] : (q[i] > bx.half[i] ? bx.half[i] : q[i]); dl[i] = q[i] - cl; if (dl[i] != 0) inside = false; }
          if (!inside) {
            const real len = N::sqrt(dot3(dl, dl));
            dist = len - r;
            for (int i = 0; i < 3; i++) nl[i] = dl[i] / len;
          } else {
            int best = 0;
            real depth = N::big;
            for (int i = 0; i < 3; i++) { const real e = bx.half[i] - N::abs(q[i]); if (e < depth) { depth = e; best = i; } }
            dist = -depth - r;
            nl[best] = q[best] >= 0 ? real(1) : real(-1);
          }
          if (dist > margin) return;
          real nw[3];
          mul_mv(nw, bx.mat, nl);
          Cand c;
          c.dist = dist; c.sign = robot_first ? real(-1) : real(1); c.wg = b;
          for (int i = 0; i < 3; i++) { c.pos[i] = p[i] - nw[i] * (r + real(0.5) * dist); c.nrm[i] = robot_first ? -nw[i] : nw[i]; }
          cand_insert(list, n, c);
        }
      }
    }
  }

  // Capsule against static boxes beyond its two end spheres: where an interior point of the axis is strictly nearer to a box than
  // both ends (a leg lying across a stair edge) that point is a third sphere feature; not generated when the ends are as near
  // (capsule flat on a face: the two end contacts carry it).  The nearest point is the root of the monotone derivative of the
  // squared segment-box distance, found by bisection.  [MJ-approx of mjc_CapsuleBox]
  QS_DEV void capsule_mid_vs_boxes(const real* c, const real* a, real L, real r, real margin, unsigned boxmask, Cand* list, int& n) const {
    unsigned mask = boxmask;
    while (mask) {
      const int k = ctz(mask);
      mask &= mask - 1;
      const DBox<real>& bx = near_box(k);
      const real rel[3] = {c[0] - bx.pos[0], c[1] - bx.pos[1], c[2] - bx.pos[2]};
      real q0[3], dv[3];
      mul_mtv(q0, bx.mat, rel);
      mul_mtv(dv, bx.mat, a);
      auto g = [&](real t) {
        real s = 0;
        for (int i = 0; i < 3; i++) {
          const real q = q0[i] + t * dv[i];
          s += q > bx.half[i] ? dv[i] * (q - bx.half[i]) : (q < -bx.half[i] ? dv[i] * (q + bx.half[i]) : real(0));
        }
        return s;
      };
      real lo = -L, hi = L;
      if (g(lo) >= 0 || g(hi) <= 0) continue;  // the nearest point is an end: the end spheres cover it
#pragma unroll 1
      for (int it = 0; it < (sizeof(real) == 4 ? 30 : 60); it++) {
        const real mid = real(0.5) * (lo + hi);
        if (g(mid) < 0) lo = mid; else hi = mid;
      }
      const real t = real(0.5) * (lo + hi);
      if (!(t > -L * (1 - real(1e-6)) && t < L * (1 - real(1e-6)))) continue;
      const real pm[3] = {c[0] + t * a[0], c[1] + t * a[1], c[2] + t * a[2]};
      const real p1[3] = {c[0] + L * a[0], c[1] + L * a[1], c[2] + L * a[2]}, p2[3] = {c[0] - L * a[0], c[1] - L * a[1], c[2] - L * a[2]};
      real nw[3];
      const real dm = point_box_distance(bx, pm, nw), d1 = point_box_distance(bx, p1, nw), d2 = point_box_distance(bx, p2, nw);
      if (!(dm < N::min(d1, d2) - real(1e-6))) continue;
      point_vs_box(pm, r, margin, true, k, list, n);
    }
  }

  // Primitive robot geoms against the perlin height field / the static boxes, as feature points (sphere centre + radius, capsule
  // end spheres, box corners) -- exact for sphere-box and plane-like cases, an approximation of the engine's capsule-box,
  // box-box and prism-based hfield routines otherwise (documented in DESIGN.md).  Appends after the floor contacts.
  // Static boxes within reach of the robot, found once per collision pass and renumbered 0 .. nnear-1 (at most 32; more would raise
  // the overflow flag).  The first KST of them are staged in shared memory -- a contact-rich pass reads every near box dozens of
  // times (per geom, per feature point), and the lift loop of a reset repeats the pass up to 100 times -- the rest stay in global.
  int nnear = 0;
  bool near_overflow = false;
  static constexpr int KST = W::KST;
  QS_DEV const DBox<real>& near_box(int k) const { return k < KST ? w.stg.box[k] : boxes[w.near_id[k]]; }
  QS_DEV void find_near_boxes() {
    int n = 0;
    for (int wd = 0; wd < 4; wd++) {
      const int b = 32 * wd + lane;
      bool near = false;
      if (b < m.nbox) {
        const real rel[3] = {boxes[b].pos[0] - w.kin.xpos[1][0], boxes[b].pos[1] - w.kin.xpos[1][1], boxes[b].pos[2] - w.kin.xpos[1][2]};
        const real reach = boxes[b].rad + m.robot_radius;
        near = dot3(rel, rel) < reach * reach;
      }
      const unsigned mask = ballot(near);
      const int pos = n + popc(mask & ((1u << lane) - 1u));
      if (near && pos < 32) w.near_id[pos] = (unsigned char)b;
      n += popc(mask);
    }
    near_overflow = n > 32;
    nnear = n > 32 ? 32 : n;
    syncwarp();
    // cooperative copy in 16-byte chunks (DBox is a multiple of 16 bytes)
    constexpr int CH = int(sizeof(DBox<real>) / 16);
    static_assert(sizeof(DBox<real>) % 16 == 0, "DBox must be a multiple of 16 bytes");
    struct alignas(16) Chunk { unsigned v[4]; };
    const int nst = nnear < KST ? nnear : KST;
    Chunk* dst = reinterpret_cast<Chunk*>(&w.stg.box[0]);
    const Chunk* src = reinterpret_cast<const Chunk*>(boxes);
    for (int i = lane; i < nst * CH; i += 32) dst[i] = src[int(w.near_id[i / CH]) * CH + i % CH];
    syncwarp();
  }
  QS_DEV void collide_terrain(int& ncon, const int gbase) {
    unsigned boxmask = nnear >= 32 ? 0xffffffffu : ((1u << nnear) - 1u);
    Cand list[4];
    int n = 0;
    const int g = gbase + lane;
    real yh[3] = {0, 0, 0};
    bool is_caps = false;
    if (geom_on(g) && m.geom_type[g] != GEOM_MESH) {
      const int b = m.geom_body[g], type = m.geom_type[g];
      real gx[3], tmp[3];
      mul_mv(tmp, w.kin.xmat[b], m.geom_pos[g]);
      for (int i = 0; i < 3; i++) gx[i] = w.kin.xpos[b][i] + tmp[i];
      const real margin = N::max(m.geom_margin[g], m.terr_margin);
      const real* sz = m.geom_size[g];
      if (ttype() == 2) {
        // per-geom cull of the warp-wide candidate set: bounding sphere of the geom against the bounding sphere of each box
        // (a superset of what the per-feature test below accepts, so the contact set is unchanged)
        const real rb = m.geom_rbound[g] + margin + real(0.01);
        {
          unsigned mask = boxmask, keep = 0;
          while (mask) {
            const int bit = ctz(mask);
            mask &= mask - 1;
            const DBox<real>& bx = near_box(bit);
            const real rel[3] = {gx[0] - bx.pos[0], gx[1] - bx.pos[1], gx[2] - bx.pos[2]};
            // geom's bounding sphere against the box grown by its radius, in the box frame: the terrain boxes are flat slabs, for
            // which the sphere-sphere test above the robot level lets through several times more candidates
            real q[3];
            mul_mtv(q, bx.mat, rel);
            if (N::abs(q[0]) <= bx.half[0] + rb && N::abs(q[1]) <= bx.half[1] + rb && N::abs(q[2]) <= bx.half[2] + rb) keep |= 1u << bit;
          }
          boxmask = keep;
        }
      }
      if (type == GEOM_SPHERE) {
        point_vs_terrain(gx, sz[0], margin, true, boxmask, list, n);
      } else if (type_on(GEOM_CAPSULE) && type == GEOM_CAPSULE) {
        real gz[3] = {m.geom_mat[g][2], m.geom_mat[g][5], m.geom_mat[g][8]}, axis[3];
        mul_mv(axis, w.kin.xmat[b], gz);
        for (int i = 0; i < 3; i++) yh[i] = axis[i];
        is_caps = true;
        for (int s = 1; s >= -1; s -= 2) {
          const real p[3] = {gx[0] + s * axis[0] * sz[1], gx[1] + s * axis[1] * sz[1], gx[2] + s * axis[2] * sz[1]};
          point_vs_terrain(p, sz[0], margin, true, boxmask, list, n);
        }
        if (ttype() == 2) capsule_mid_vs_boxes(gx, axis, sz[1], sz[0], margin, boxmask, list, n);
      } else if ((type_on(GEOM_BOX) && type == GEOM_BOX) || (type_on(GEOM_CYLINDER) && type == GEOM_CYLINDER)) {
        // box corners / eight cylinder rim points (four per cap) as point features
        real gm[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) gm[3 * r + c] = w.kin.xmat[b][3 * r] * m.geom_mat[g][c] + w.kin.xmat[b][3 * r + 1] * m.geom_mat[g][3 + c] + w.kin.xmat[b][3 * r + 2] * m.geom_mat[g][6 + c];
        const bool cyl = type == GEOM_CYLINDER;
        for (int i = 0; i < 8; i++) {
          real v[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]};
          if (cyl) { const real a = v[0]; v[0] = (i & 2) ? a : real(0); v[1] = (i & 2) ? real(0) : a; v[2] = (i & 4) ? sz[1] : -sz[1]; }
          real corner[3];
          mul_mv(corner, gm, v);
          const real p[3] = {corner[0] + gx[0], corner[1] + gx[1], corner[2] + gx[2]};
          point_vs_terrain(p, real(0), margin, cyl, boxmask, list, n);
        }
      }
    }
    for (int r = 0; r < 4; r++) {
      const bool has = n > r;
      const unsigned mask = ballot(has);
      if (mask == 0) break;
      const int slot = ncon + popc(mask & ((1u << lane) - 1u));
      if (has) note_contact(g, list[r].dist);
      if (has && slot < NCON) store_contact(slot, g, list[r].sign, list[r].dist, list[r].pos, list[r].nrm, is_caps ? yh : nullptr, list[r].wg);
      ncon += popc(mask);
    }
  }

  // signed distance of point p to static box bx (negative inside) and the outward box normal there (world frame)
  QS_DEV static real point_box_distance(const DBox<real>& bx, const real* p, real* nw) {
    const real rel[3] = {p[0] - bx.pos[0], p[1] - bx.pos[1], p[2] - bx.pos[2]};
    real q[3], dl[3], nl[3] = {0, 0, 0}, dist;
    mul_mtv(q, bx.mat, rel);
    bool inside = true;
    for (int i = 0; i < 3; i++) { const real cl = q[i] < -bx.half[i] ? -bx.half[i] : (q[i] > bx.half[i] ? bx.half[i] : q[i]); dl[i] = q[i] - cl; if (dl[i] != 0) inside = false; }
    if (!inside) {
      const real len = N::sqrt(dot3(dl, dl));
      dist = len;
      for (int i = 0; i < 3; i++) nl[i] = dl[i] / len;
    } else {
      int best = 0;
      real depth = N::big;
      for (int i = 0; i < 3; i++) { const real e = bx.half[i] - N::abs(q[i]); if (e < depth) { depth = e; best = i; } }
      dist = -depth;
      nl[best] = q[best] >= 0 ? real(1) : real(-1);
    }
    mul_mv(nw, bx.mat, nl);
    return dist;
  }

  // Convex meshes against the height field / the static boxes, hull vertices as feature points (the terrain counterpart of the
  // support-vertex rule used on the floor plane): the deepest vertex per mesh on the height field, the deepest vertex per
  // (mesh, box) pair on boxes, at most 4 per mesh, deepest first.  Broad phase on the lanes (bounding sphere of the hull's box
  // against a Lipschitz bound of the field around it / against the robot-near boxes), the vertex scans by the whole warp.
  QS_DEV void collide_mesh_terrain(const int gbase, int& ncon) {
    const int g = gbase + lane;
    const int tt = ttype();
    bool near = false;
    if (geom_on(g) && m.geom_type[g] == GEOM_MESH) {
      const int b = m.geom_body[g];
      real c[3], tmp[3];
      mul_mv(tmp, w.kin.xmat[b], m.geom_bcenter[g]);
      for (int i = 0; i < 3; i++) c[i] = w.kin.xpos[b][i] + tmp[i];
      const real* bh = m.geom_bhalf[g];
      const real rb = N::sqrt(bh[0] * bh[0] + bh[1] * bh[1] + bh[2] * bh[2]);
      if (tt == 1) {
        // the field under the hull's footprint is at most lip * rb above its height at the centre (anywhere: below the global top)
        real z, nn[3];
        const real top = m.hf_pos[2] + m.hf_size[2];
        const real bound = hfield_height(c[0], c[1], z, nn) ? N::min(top, z + m.hf_lip * rb) : top;
        const bool over = !(c[0] + rb < m.hf_pos[0] - m.hf_size[0] || c[0] - rb > m.hf_pos[0] + m.hf_size[0] ||
                            c[1] + rb < m.hf_pos[1] - m.hf_size[1] || c[1] - rb > m.hf_pos[1] + m.hf_size[1]);
        near = over && !(c[2] - rb > bound + N::max(m.geom_margin[g], m.terr_margin) + real(1e-4));
      } else {
        near = nnear > 0;
      }
    }
    unsigned cand = ballot(near);
    while (cand) {
      const int gm_ = gbase + ctz(cand);
      cand &= cand - 1;
      const int b = m.geom_body[gm_];
      const real margin = N::max(m.geom_margin[gm_], m.terr_margin);
      const real* R = w.kin.xmat[b];
      const real* X = w.kin.xpos[b];
      const Vert4<real>* v = vert + m.geom_vertadr[gm_];
      const int nv = m.geom_vertnum[gm_];
      if (tt == 1) {
        real best = N::big;
        int bi = 0x7fffffff;
        for (int i = lane; i < nv; i += 32) {
          const Vert4<real> q = v[i];
          const real p[3] = {X[0] + R[0] * q.x + R[1] * q.y + R[2] * q.z, X[1] + R[3] * q.x + R[4] * q.y + R[5] * q.z, X[2] + R[6] * q.x + R[7] * q.y + R[8] * q.z};
          real z, nn[3];
          if (!hfield_height(p[0], p[1], z, nn)) continue;
          const real dist = (p[2] - z) * nn[2];
          if (dist < best) { best = dist; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const real ob = shfl_xor(best, o);
          const int oi = shfl_xor(bi, o);
          if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (bi == 0x7fffffff || best > margin) continue;
        note_contact(gm_, best);
        if (lane == 0 && ncon < NCON) {
          const Vert4<real> q = v[bi];
          const real p[3] = {X[0] + R[0] * q.x + R[1] * q.y + R[2] * q.z, X[1] + R[3] * q.x + R[4] * q.y + R[5] * q.z, X[2] + R[6] * q.x + R[7] * q.y + R[8] * q.z};
          real z, nn[3];
          hfield_height(p[0], p[1], z, nn);
          const real pos[3] = {p[0] - nn[0] * real(0.5) * best, p[1] - nn[1] * real(0.5) * best, p[2] - nn[2] * real(0.5) * best};
          store_contact(ncon, gm_, real(1), best, pos, nn, nullptr, WG_HFIELD);
        }
        ncon++;
      } else {
        // every lane keeps the same (warp-uniform) list of the 4 deepest (box, vertex) pairs of this mesh
        real ld[4];
        int lb[4], lv[4], n = 0;
        real c[3], tmp[3];
        mul_mv(tmp, R, m.geom_bcenter[gm_]);
        for (int i = 0; i < 3; i++) c[i] = X[i] + tmp[i];
        const real* bh = m.geom_bhalf[gm_];
        const real rb = N::sqrt(bh[0] * bh[0] + bh[1] * bh[1] + bh[2] * bh[2]) + margin + real(0.01);
        {
          for (int kb = 0; kb < nnear; kb++) {
            const int bx = w.near_id[kb];
            const DBox<real>& B = near_box(kb);
            const real rel[3] = {c[0] - B.pos[0], c[1] - B.pos[1], c[2] - B.pos[2]};
            real qc[3];
            mul_mtv(qc, B.mat, rel);
            if (N::abs(qc[0]) > B.half[0] + rb || N::abs(qc[1]) > B.half[1] + rb || N::abs(qc[2]) > B.half[2] + rb) continue;
            real best = N::big;
            int bi = 0x7fffffff;
            for (int i = lane; i < nv; i += 32) {
              const Vert4<real> q = v[i];
              const real p[3] = {X[0] + R[0] * q.x + R[1] * q.y + R[2] * q.z, X[1] + R[3] * q.x + R[4] * q.y + R[5] * q.z, X[2] + R[6] * q.x + R[7] * q.y + R[8] * q.z};
              real nw[3];
              const real dist = point_box_distance(B, p, nw);
              if (dist < best) { best = dist; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
              const real ob = shfl_xor(best, o);
              const int oi = shfl_xor(bi, o);
              if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (best > margin) continue;
            // insertion by depth; among equals the earlier box stays first (the order of a stable sort over the box index)
            int k = n < 4 ? n : 4;
            if (n >= 4 && !(best < ld[3])) continue;
            if (k == 4) k = 3;
            while (k > 0 && best < ld[k - 1]) { ld[k] = ld[k - 1]; lb[k] = lb[k - 1]; lv[k] = lv[k - 1]; k--; }
            ld[k] = best; lb[k] = bx; lv[k] = bi;
            if (n < 4) n++;
          }
        }
        for (int r = 0; r < n; r++) {
          note_contact(gm_, ld[r]);
          if (lane == 0 && ncon < NCON) {
            const Vert4<real> q = v[lv[r]];
            const real p[3] = {X[0] + R[0] * q.x + R[1] * q.y + R[2] * q.z, X[1] + R[3] * q.x + R[4] * q.y + R[5] * q.z, X[2] + R[6] * q.x + R[7] * q.y + R[8] * q.z};
            real nw[3];
            point_box_distance(boxes[lb[r]], p, nw);
            const real pos[3] = {p[0] - nw[0] * real(0.5) * ld[r], p[1] - nw[1] * real(0.5) * ld[r], p[2] - nw[2] * real(0.5) * ld[r]};
            store_contact(ncon, gm_, real(1), ld[r], pos, nw, nullptr, lb[r]);
          }
          ncon++;
        }
      }
    }
  }

  // downward ray from `org` against the static terrain: distance to the nearest hit or -1 ([MJ] mj_ray, heightmap.py:77-99)
  QS_DEV real ray_down(const real* org) const {
    real best = org[2] >= 0 ? org[2] : real(-1);  // floor plane z = 0
    if (!terrain_on) return best;
    if (ttype() == 1) {
      real z, nn[3];
      if (hfield_height(org[0], org[1], z, nn) && org[2] >= z) { const real t = org[2] - z; if (best < 0 || t < best) best = t; }
    } else if (ttype() == 2) {
      for (int b = 0; b < m.nbox; b++) {
        const DBox<real>& bx = boxes[b];
        const real rel[3] = {org[0] - bx.pos[0], org[1] - bx.pos[1], org[2] - bx.pos[2]};
        if (rel[0] * rel[0] + rel[1] * rel[1] > bx.rad * bx.rad) continue;
        real o[3];
        mul_mtv(o, bx.mat, rel);
        const real dl[3] = {-bx.mat[6], -bx.mat[7], -bx.mat[8]};  // R^T (0,0,-1)
        real tmin = -N::big, tmax = N::big;
        bool miss = false;
        for (int i = 0; i < 3; i++) {
          if (N::abs(dl[i]) < real(1e-12)) { if (o[i] < -bx.half[i] || o[i] > bx.half[i]) miss = true; continue; }
          real t1 = (-bx.half[i] - o[i]) / dl[i], t2 = (bx.half[i] - o[i]) / dl[i];
          if (t1 > t2) { const real t = t1; t1 = t2; t2 = t; }
          tmin = N::max(tmin, t1); tmax = N::min(tmax, t2);
        }
        if (miss || tmin > tmax || tmax < 0) continue;
        const real t = tmin >= 0 ? tmin : tmax;
        if (best < 0 || t < best) best = t;
      }
    }
    return best;
  }

  // HeightMap.create_sensor_matrix (sensors/heightmap.py:106-169): rows x cols hit points around `center` with heading `yaw`,
  // written as [rows][cols][3] floats to `out` (global memory); lanes stride over the grid cells
  QS_DEV void heightmap(const real* center, real yaw, int rows, int cols, real dx, real dy, real ox, real oy, float* out) const {
    const real c_rows = rows % 2 == 0 ? real(rows) / 2 : real(rows - 1) / 2, add_r = rows % 2 == 0 ? -dx / 2 : real(0);
    const real c_cols = cols % 2 == 0 ? real(cols) / 2 : real(cols - 1) / 2, add_c = cols % 2 == 0 ? -dy / 2 : real(0);
    real sy, cy;
    N::sincos(yaw, &sy, &cy);
    for (int it = lane; it < rows * cols; it += 32) {
      const int i = it / cols, j = it % cols;
      const real offx = dx * (c_rows - i) + add_r, offy = dy * (c_cols - j) + add_c;
      const real org[3] = {center[0] + cy * offx - sy * offy, center[1] + sy * offx + cy * offy, center[2] + real(0.6) - real(0.07)};
      const real t = ray_down(org);
      out[3 * it] = float(org[0] + ox); out[3 * it + 1] = float(org[1] + oy); out[3 * it + 2] = float(org[2] - t);
    }
  }

  // floor plane z = 0 (scene_flat.xml:32) against every robot geom. [MJ] mjc_PlaneSphere/Capsule/Box/Convex
  QS_DEV void collide_floor() {
    int ncon = 0;
    cm_acc = im_acc = 0; pen_acc = 0;
    near_overflow = false;  // a reset pass may land outside the terrain patch: nothing of the previous pass's box search carries over
    if (ttype() == 2 && terrain_on) find_near_boxes();
    // one lane per geom; robots with more than 32 collision geoms (go1: 42) take a second round
    if (FEAT & FEAT_NGEOM32) collide_round(0, ncon);
    else {
#pragma unroll 1
      for (int gbase = 0; gbase < m.ngeom; gbase += 32) collide_round(gbase, ncon);
    }
    if (lane == 0) { w.overflow = ncon > NCON || near_overflow; w.ncon = ncon > NCON ? NCON : ncon; }
    syncwarp();
  }
  QS_DEV void collide_round(const int gbase, int& ncon) {
    // primitives: one lane per geom, up to 4 candidate contacts each
    real cd[4], cp[4][3], yh[3] = {0, 0, 0};
    int nc = 0;
    const int g = gbase + lane;
    if (geom_on(g) && m.geom_type[g] != GEOM_MESH) {
      const int b = m.geom_body[g];
      real gx[3], tmp[3];
      mul_mv(tmp, w.kin.xmat[b], m.geom_pos[g]);
      for (int i = 0; i < 3; i++) gx[i] = w.kin.xpos[b][i] + tmp[i];
      const real margin = m.geom_margin[g];
      const real* sz = m.geom_size[g];
      const int type = m.geom_type[g];
      if (m.geom_leg[g] >= 0) for (int i = 0; i < 3; i++) w.footpos[m.geom_leg[g]][i] = gx[i];
      if (type == GEOM_SPHERE) {
        real dist = gx[2] - sz[0];
        if (!(dist > margin)) { cd[0] = dist; cp[0][0] = gx[0]; cp[0][1] = gx[1]; cp[0][2] = gx[2] - (sz[0] + real(0.5) * dist); nc = 1; }
      } else if (type_on(GEOM_CAPSULE) && type == GEOM_CAPSULE) {
        // capsule axis = third column of (R_body * R_geom)
        real gz[3] = {m.geom_mat[g][2], m.geom_mat[g][5], m.geom_mat[g][8]}, axis[3];
        mul_mv(axis, w.kin.xmat[b], gz);
        for (int i = 0; i < 3; i++) yh[i] = axis[i];
        for (int s = 1; s >= -1; s -= 2) {
          real p[3] = {gx[0] + s * axis[0] * sz[1], gx[1] + s * axis[1] * sz[1], gx[2] + s * axis[2] * sz[1]};
          real dist = p[2] - sz[0];
          if (dist > margin) continue;
          cd[nc] = dist; cp[nc][0] = p[0]; cp[nc][1] = p[1]; cp[nc][2] = p[2] - (sz[0] + real(0.5) * dist); nc++;
        }
      } else if (type_on(GEOM_CYLINDER) && type == GEOM_CYLINDER) {
        // [MJ] mjc_PlaneCylinder: nearest rim point of the lower cap, its twin on the other cap, two more lower-rim points at +-120 deg
        real gxa[3] = {m.geom_mat[g][0], m.geom_mat[g][3], m.geom_mat[g][6]}, gza[3] = {m.geom_mat[g][2], m.geom_mat[g][5], m.geom_mat[g][8]};
        real xax[3], axis[3], rim[3];
        mul_mv(xax, w.kin.xmat[b], gxa);
        mul_mv(axis, w.kin.xmat[b], gza);
        real axis_z = axis[2];
        if (axis_z > 0) { for (int i = 0; i < 3; i++) axis[i] = -axis[i]; axis_z = -axis_z; }
        const real height = gx[2];
        rim[0] = axis[0] * axis_z; rim[1] = axis[1] * axis_z; rim[2] = axis[2] * axis_z - 1;
        const real rim_n2 = dot3(rim, rim);
        if (rim_n2 >= real(1e-30)) { const real scl = sz[0] * N::rsqrt(rim_n2); for (int i = 0; i < 3; i++) rim[i] *= scl; }
        else for (int i = 0; i < 3; i++) rim[i] = xax[i] * sz[0];
        const real rim_z = rim[2];
        const real half[3] = {axis[0] * sz[1], axis[1] * sz[1], axis[2] * sz[1]};
        axis_z *= sz[1];
        real dist = height + axis_z + rim_z;
        if (!(dist > margin)) {
          cd[0] = dist; for (int i = 0; i < 3; i++) cp[0][i] = gx[i] + rim[i] + half[i]; cp[0][2] -= real(0.5) * dist; nc = 1;
          dist = height - axis_z + rim_z;
          if (!(dist > margin)) { cd[nc] = dist; for (int i = 0; i < 3; i++) cp[nc][i] = gx[i] + rim[i] - half[i]; cp[nc][2] -= real(0.5) * dist; nc++; }
          dist = height + axis_z - real(0.5) * rim_z;
          if (!(dist > margin)) {
            real side[3];
            cross3(side, rim, half);
            const real n2 = dot3(side, side), scl = n2 > 0 ? sz[0] * real(0.86602540378443864676) * N::rsqrt(n2) : real(0);
            for (int sgn = 1; sgn >= -1; sgn -= 2) {
              cd[nc] = dist;
              for (int i = 0; i < 3; i++) cp[nc][i] = gx[i] + sgn * scl * side[i] + half[i] - real(0.5) * rim[i];
              cp[nc][2] -= real(0.5) * dist;
              nc++;
            }
          }
        }
      } else if (type_on(GEOM_BOX) && type == GEOM_BOX) {
        real gm[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) gm[3 * r + c] = w.kin.xmat[b][3 * r] * m.geom_mat[g][c] + w.kin.xmat[b][3 * r + 1] * m.geom_mat[g][3 + c] + w.kin.xmat[b][3 * r + 2] * m.geom_mat[g][6 + c];
        for (int i = 0; i < 8; i++) {
          if (nc >= 4) break;
          real v[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]}, corner[3];
          mul_mv(corner, gm, v);
          real ldist = corner[2];
          if (gx[2] + ldist > margin || ldist > 0) continue;
          real dist = gx[2] + ldist;
          cd[nc] = dist; cp[nc][0] = corner[0] + gx[0]; cp[nc][1] = corner[1] + gx[1]; cp[nc][2] = corner[2] + gx[2] - real(0.5) * dist; nc++;
        }
      }
    }
    const real nrm[3] = {0, 0, 1};
    for (int r = 0; r < 4; r++) {
      const bool has = nc > r;
      const unsigned mask = ballot(has);
      if (mask == 0) break;
      const int slot = ncon + popc(mask & ((1u << lane) - 1u));
      if (has) {
        note_contact(g, cd[r]);
        if (slot < NCON) store_contact(slot, g, real(1), cd[r], cp[r], nrm, (m.geom_type[g] == GEOM_CAPSULE) ? yh : nullptr, WG_FLOOR);
      }
      ncon += popc(mask);
    }
    if (ttype() != 0 && terrain_on) collide_terrain(ncon, gbase);
    // convex meshes. Broad phase: lanes test the body-frame bounding box of every mesh against the plane (a lower bound of the
    // hull's lowest point, tight for long thin links); only the survivors are scanned, by the whole warp, for their support vertex.
    unsigned cand = 0;
    if (type_on(GEOM_MESH)) {
      bool near = false;
      if (type_on(GEOM_MESH) && geom_on(g) && m.geom_type[g] == GEOM_MESH) {
        const int b = m.geom_body[g];
        const real* R = w.kin.xmat[b];
        const real cz = w.kin.xpos[b][2] + R[6] * m.geom_bcenter[g][0] + R[7] * m.geom_bcenter[g][1] + R[8] * m.geom_bcenter[g][2];
        const real ext = N::abs(R[6]) * m.geom_bhalf[g][0] + N::abs(R[7]) * m.geom_bhalf[g][1] + N::abs(R[8]) * m.geom_bhalf[g][2];
        near = !(cz - ext > m.geom_margin[g] + real(1e-5));  // slack: the box is only a filter, never the decision
      }
      cand = ballot(near);
    }
    while (cand) {
      const int gm_ = gbase + ctz(cand);
      cand &= cand - 1;
      const int b = m.geom_body[gm_];
      const real margin = m.geom_margin[gm_];
      const real* R = w.kin.xmat[b];
      const real dx = -R[6], dy = -R[7], dz = -R[8];  // R^T * (-normal)
      const Vert4<real>* v = vert + m.geom_vertadr[gm_];
      const int nv = m.geom_vertnum[gm_];
      real best = -N::big;
      int bi = 0x7fffffff;
      for (int i = lane; i < nv; i += 32) {
        const Vert4<real> p = v[i];
        const real s = p.x * dx + p.y * dy + p.z * dz;
        if (s > best) { best = s; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const real ob = shfl_xor(best, o);
        const int oi = shfl_xor(bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (bi == 0x7fffffff) continue;  // non-finite pose: no vertex compares greater (the step then ends the episode, status bit0)
      const Vert4<real> p = v[bi];
      const real pw[3] = {w.kin.xpos[b][0] + R[0] * p.x + R[1] * p.y + R[2] * p.z, w.kin.xpos[b][1] + R[3] * p.x + R[4] * p.y + R[5] * p.z,
                          w.kin.xpos[b][2] + R[6] * p.x + R[7] * p.y + R[8] * p.z};
      const real dist = pw[2];
      if (dist > margin) continue;
      note_contact(gm_, dist);
      if (lane == 0 && ncon < NCON) {
        const real pos[3] = {pw[0], pw[1], pw[2] - real(0.5) * dist};
        store_contact(ncon, gm_, real(1), dist, pos, nrm, nullptr, WG_FLOOR);
      }
      ncon++;
    }
    if (type_on(GEOM_MESH) && ttype() != 0 && terrain_on) collide_mesh_terrain(gbase, ncon);
  }

  // ------------------------------------------------------------------ constraint construction
  // [MJ] getimpedance
  QS_DEV real impedance(const real* si, real x) const {
    real d0 = N::min(N::max(si[0], real(0.0001)), real(0.9999)), d1 = N::min(N::max(si[1], real(0.0001)), real(0.9999));
    real width = N::max(si[2], real(0)), mid = N::min(N::max(si[3], real(0.0001)), real(0.9999)), power = N::max(si[4], real(1));
    if (d0 == d1 || width <= N::minval) return real(0.5) * (d0 + d1);
    x = N::abs(x) / width;
    if (x >= 1) return d1;
    if (x == 0) return d0;
    real y;
    if (power == real(1)) y = x;
    else if (power == real(2)) y = (x <= mid) ? x * x / mid : real(1) - (real(1) - x) * (real(1) - x) / (real(1) - mid);
    else if (x <= mid) y = N::pow(x, power) / N::pow(mid, power - 1);
    else y = real(1) - N::pow(real(1) - x, power) / N::pow(real(1) - mid, power - 1);
    return d0 + y * (d1 - d0);
  }

  // generalized-velocity index of contact-Jacobian column col (0..5 base, 6..8 leg of the contact body)
  QS_DEV static int col_dof(int body, int col) { return col < 6 ? col : 6 + 3 * ((body - 2) / 3) + (col - 6); }

  // [MJ] mj_makeConstraint / mj_makeImpedance / mj_referenceConstraint in "unit" form:
  //   friction-loss and limit units are scalar; a contact unit carries its contact-frame Jacobian Jc (dim x 9), the
  //   contact-frame reference acceleration ar and regularisers D; pyramid rows are expanded on the fly as r0 +- mu r_j.
  QS_DEV void make_constraints() {
    const int ncon = w.ncon;
    // contact Jacobians: item (c, col)
    for (int it = lane; it < ncon * 9; it += 32) {
      const int c = it / 9, col = it % 9, info = w.c_info[c], body = info_body(info), dim = info_dim(info);
      bool on = col < 6;
      if (!on && body >= 2) on = (col - 6) <= (body - 2) % 3;
      real jp[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
      if (on) {
        const int d = col_dof(body, col);
        const real off[3] = {w.c_pos[c][0] - w.com[0], w.c_pos[c][1] - w.com[1], w.c_pos[c][2] - w.com[2]};
        real cr[3];
        cross3(cr, w.cdof[d], off);
        for (int i = 0; i < 3; i++) { jp[i] = w.cdof[d][3 + i] + cr[i]; jr[i] = w.cdof[d][i]; }
      }
      const real sg = w.c_sign[c];
      for (int k = 0; k < 3; k++) {
        real fr[3];
        frame_row(c, k, fr);
        // rows beyond the contact's dimension are written as zeros so that the solver loop can run over all MAXDIM rows unguarded
        w.Jc[c][k][col] = (k < dim) ? sg * dot3(fr, jp) : real(0);
        if (MAXDIM > 3) w.Jc[c][MAXDIM > 3 ? 3 + k : 0][col] = (3 + k < dim) ? sg * dot3(fr, jr) : real(0);
      }
    }
    // scalar units
    if (lane < NFL) {
      const int d = 6 + lane;
      w.u_D[lane] = m.dof_floss[d] > 0 ? m.dof_D[d] : real(0);
      w.u_sign[lane] = 1;
      w.u_ar[lane] = -m.dof_B[d] * w.qvel[d];
    } else if (kLimits && lane < NFL + NLIM) {
      const int j = lane - NFL;
      real D = 0, ar = 0, sign = 0;
      if (m.jnt_limited[j]) {
        const real value = w.qpos[7 + j], dlo = value - m.jnt_range[j][0], dhi = m.jnt_range[j][1] - value, margin = m.jnt_margin[j];
        real pos = 0;
        if (dlo < margin) { sign = 1; pos = dlo; } else if (dhi < margin) { sign = -1; pos = dhi; }
        if (sign != 0) {
          const real imp = impedance(m.jnt_solimp[j], pos - margin);
          D = 1 / N::max(N::minval, (1 - imp) * m.dof_iw[6 + j] / imp);
          ar = -m.jnt_B[j] * (sign * w.qvel[6 + j]) - m.jnt_K[j] * imp * (pos - margin);
        }
      }
      w.u_D[lane] = D; w.u_sign[lane] = sign; w.u_ar[lane] = ar;
    }
    syncwarp();
    for (int c = lane; c < ncon; c += 32) {
      const int info = w.c_info[c], g = info_geom(info), body = info_body(info), dim = info_dim(info);
      real velc[MAXDIM];
      for (int k = 0; k < MAXDIM; k++) {
        real s = 0;
        if (k < dim) {
          const real* J = w.Jc[c][k];
          s = J[0] * w.qvel[0] + J[1] * w.qvel[1] + J[2] * w.qvel[2] + J[3] * w.qvel[3] + J[4] * w.qvel[4] + J[5] * w.qvel[5];
          if (body >= 2) { const real* xl = w.qvel + 6 + 3 * info_leg(info); s += J[6] * xl[0] + J[7] * xl[1] + J[8] * xl[2]; }
        }
        velc[k] = s;
      }
      const real x = w.c_dist[c] - m.geom_incmargin[g];
      const bool active = w.c_dist[c] < m.geom_incmargin[g];
      const bool tw = info_terrain_wins(info);
      const real imp = impedance(tw ? m.terr_solimp : m.geom_solimp[g], x);
      const real tran = m.body_iw[body][0];
      const real K = tw ? m.terr_K : m.geom_K[g], B = tw ? m.terr_B : m.geom_B[g];
      const real f0 = w.c_fri[c][0];
      for (int k = 0; k < MAXDIM; k++) { w.c_D[c][k] = 0; w.c_ar[c][k] = -B * velc[k]; }
      w.c_ar[c][0] -= K * imp * x;
      if (!active) { w.c_mu[c] = f0; continue; }
      if (dim == 1) {
        w.c_D[c][0] = 1 / N::max(N::minval, (1 - imp) * tran / imp);
        w.c_mu[c] = f0;
      } else if (cone_is_pyramidal()) {
        const real Rn = N::max(N::minval, (1 - imp) * (tran + f0 * f0 * tran) / imp);
        const real mur = f0 * N::sqrt(1 / N::max(N::minval, m.impratio));
        const real Rpy = 2 * mur * mur * Rn;
        for (int k = 0; k < MAXDIM; k++) w.c_D[c][k] = 1 / Rpy;
        w.c_mu[c] = f0;  // pyramid rows use the friction coefficient itself
      } else {
        const real R0 = N::max(N::minval, (1 - imp) * tran / imp), R1 = R0 / N::max(N::minval, m.impratio);
        w.c_mu[c] = f0 * N::sqrt(R1 / R0);
        w.c_D[c][0] = 1 / R0; w.c_D[c][1] = 1 / R1; w.c_D[c][2] = 1 / R1;
        if (MAXDIM > 3 && dim > 3) {
          const real f1 = w.c_fri[c][1], f2 = w.c_fri[c][2];
          w.c_D[c][3] = 1 / (R1 * f0 * f0 / (f1 * f1));
          w.c_D[c][MAXDIM > 4 ? 4 : 0] = 1 / (R1 * f0 * f0 / (f2 * f2));
          w.c_D[c][MAXDIM > 5 ? 5 : 0] = 1 / (R1 * f0 * f0 / (f2 * f2));
        }
      }
    }
    syncwarp();
  }

  // friction coefficient that multiplies contact-frame dimension k (k>=1): slide, slide, spin, roll, roll
  QS_DEV real fri_k(int c, int k) const { return k <= 2 ? w.c_fri[c][0] : (k == 3 ? w.c_fri[c][1] : w.c_fri[c][2]); }

  // ------------------------------------------------------------------ per-unit cost model  [MJ] mj_constraintUpdate
  // branch-free (selects): these sit in the innermost loop of the line search, where every divergent region costs a re-convergence
  QS_DEV static void row_q(real x, real v, real D, real& cost, real& d1, real& d2) {
    const bool on = x < 0;
    const real c = real(0.5) * D * x * x, a = D * x * v, b = D * v * v;
    cost += on ? c : real(0); d1 += on ? a : real(0); d2 += on ? b : real(0);
  }
  // scalar unit u at residual x with direction v
  QS_DEV void scalar_unit_eval(int u, real x, real v, real& cost, real& d1, real& d2) const {
    const real D = w.u_D[u];
    const bool fric = u < NFL;
    const real f = m.dof_floss[6 + (fric ? u : 0)], rf = N::div(f, D != 0 ? D : real(1));
    const bool lo = fric && x <= -rf, hi = fric && !lo && x >= rf;       // linear zones of the friction-loss cost
    const bool quad = D != 0 && (fric ? !(lo || hi) : x < 0);
    const real cq = real(0.5) * D * x * x, aq = D * x * v, bq = D * v * v;
    const real cl = -f * (real(0.5) * rf + x), ch = -f * (real(0.5) * rf - x);
    const bool act = D != 0;
    cost += quad ? cq : ((act && lo) ? cl : ((act && hi) ? ch : real(0)));
    d1 += quad ? aq : ((act && lo) ? -f * v : ((act && hi) ? f * v : real(0)));
    d2 += quad ? bq : real(0);
  }
  // contact unit c at contact-frame residual r[] with direction v[]
  QS_DEV void contact_unit_eval(int c, const real* r, const real* v, real& cost, real& d1, real& d2) const {
    const int dim = info_dim(w.c_info[c]);
    const real D0 = w.c_D[c][0];
    if (D0 == 0) return;
    if (dim == 1) { row_q(r[0], v[0], D0, cost, d1, d2); return; }
    const real mu = w.c_mu[c];
    if (cone_is_pyramidal()) {
      for (int j = 1; j < 3; j++) {
        row_q(r[0] + mu * r[j], v[0] + mu * v[j], D0, cost, d1, d2);
        row_q(r[0] - mu * r[j], v[0] - mu * v[j], D0, cost, d1, d2);
      }
      return;
    }
    const real Nn = r[0] * mu, N1 = v[0] * mu;
    real T2 = 0, UV = 0, VV = 0;
    for (int k = 1; k < MAXDIM; k++) if (k < dim) { const real f = fri_k(c, k), U = r[k] * f, V = v[k] * f; T2 += U * U; UV += U * V; VV += V * V; }
    const real T = N::sqrt(T2);
    if (Nn >= mu * T || (T <= 0 && Nn >= 0)) return;
    if (mu * Nn + T <= 0 || (T <= 0 && Nn < 0)) {
      for (int k = 0; k < MAXDIM; k++) if (k < dim) { const real Dk = w.c_D[c][k]; cost += real(0.5) * Dk * r[k] * r[k]; d1 += Dk * r[k] * v[k]; d2 += Dk * v[k] * v[k]; }
      return;
    }
    const real Dm = N::div(D0, N::max(N::minval, mu * mu * (1 + mu * mu))), NmT = Nn - mu * T;
    const real invT = N::rcp(T);  // one reciprocal instead of three divisions (innermost loop of the line search)
    const real T1 = UV * invT, T2d = VV * invT - UV * UV * (invT * invT * invT), e1 = N1 - mu * T1;
    cost += real(0.5) * Dm * NmT * NmT;
    d1 += Dm * NmT * e1;
    d2 += Dm * (e1 * e1 - NmT * mu * T2d);
  }
  // contact unit c: contact-frame force F and packed-symmetric Hessian weight Wt at residual r; returns cost
  QS_DEV real contact_unit_update(int c) {
    const int dim = info_dim(w.c_info[c]);
    const real* r = w.c_r[c];
    real* F = w.c_F[c];
    real* Wt = w.c_W[c];
    for (int a = 0; a < MAXDIM; a++) F[a] = 0;
    for (int a = 0; a < NW; a++) Wt[a] = 0;
    const real D0 = w.c_D[c][0];
    if (D0 == 0) return 0;
    real cost = 0;
    if (dim == 1) {
      if (r[0] < 0) { F[0] = -D0 * r[0]; Wt[0] = D0; cost = real(0.5) * D0 * r[0] * r[0]; }
      return cost;
    }
    const real mu = w.c_mu[c];
    if (cone_is_pyramidal()) {
      real fe[4], de[4];
      for (int e = 0; e < 4; e++) {
        const real x = r[0] + ((e & 1) ? -mu : mu) * r[1 + e / 2];
        const bool act = x < 0;
        fe[e] = act ? -D0 * x : real(0);
        de[e] = act ? D0 : real(0);
        if (act) cost += real(0.5) * D0 * x * x;
      }
      F[0] = fe[0] + fe[1] + fe[2] + fe[3]; F[1] = mu * (fe[0] - fe[1]); F[2] = mu * (fe[2] - fe[3]);
      Wt[widx(0, 0)] = de[0] + de[1] + de[2] + de[3];
      Wt[widx(0, 1)] = mu * (de[0] - de[1]);
      Wt[widx(0, 2)] = mu * (de[2] - de[3]);
      Wt[widx(1, 1)] = mu * mu * (de[0] + de[1]);
      Wt[widx(2, 2)] = mu * mu * (de[2] + de[3]);
      return cost;
    }
    real U[MAXDIM], fk[MAXDIM];
    U[0] = r[0] * mu; fk[0] = mu;
    real T2 = 0;
    for (int k = 1; k < MAXDIM; k++) { fk[k] = (k < dim) ? fri_k(c, k) : real(0); U[k] = (k < dim) ? r[k] * fk[k] : real(0); T2 += U[k] * U[k]; }
    const real Nn = U[0], T = N::sqrt(T2);
    if (Nn >= mu * T || (T <= 0 && Nn >= 0)) return 0;
    if (mu * Nn + T <= 0 || (T <= 0 && Nn < 0)) {
      for (int k = 0; k < MAXDIM; k++) if (k < dim) { const real Dk = w.c_D[c][k]; F[k] = -Dk * r[k]; Wt[widx(k, k)] = Dk; cost += real(0.5) * Dk * r[k] * r[k]; }
      return cost;
    }
    const real Dm = N::div(D0, N::max(N::minval, mu * mu * (1 + mu * mu))), NmT = Nn - mu * T;
    cost = real(0.5) * Dm * NmT * NmT;
    F[0] = -Dm * NmT * mu;
    real de[MAXDIM];
    de[0] = mu;
    const real invT = N::rcp(T), invT3 = invT * invT * invT, c1 = Dm * NmT * (-mu);  // reciprocals once, not per Hessian entry
    for (int k = 1; k < MAXDIM; k++) { de[k] = (k < dim) ? -mu * fk[k] * U[k] * invT : real(0); if (k < dim) F[k] = -F[0] * invT * U[k] * fk[k]; }
    for (int a = 0; a < MAXDIM; a++)
      for (int b = a; b < MAXDIM; b++) {
        if (a >= dim || b >= dim) continue;
        real h = Dm * de[a] * de[b];
        if (a > 0) h += c1 * fk[a] * fk[b] * ((a == b ? invT : real(0)) - U[a] * U[b] * invT3);
        Wt[widx(a, b)] = h;
      }
    return cost;
  }

  // J x for every unit: out_u[] (scalar units) and out_c[][] (contacts). x in shared memory. minus_ar: subtract reference.
  QS_DEV void units_Jx(const real* x, real* out_u, real (*out_c)[MAXDIM], bool minus_ar) {
    if (lane < NSC) {
      const int d = 6 + (lane < NFL ? lane : lane - NFL);
      out_u[lane] = w.u_sign[lane] * x[d] - (minus_ar ? w.u_ar[lane] : real(0));
    }
    const int ncon = w.ncon;
    for (int it = lane; it < ncon * MAXDIM; it += 32) {
      const int c = it / MAXDIM, k = it % MAXDIM, info = w.c_info[c], body = info_body(info);
      const real* J = w.Jc[c][k];  // rows >= dim hold zeros (and a zero reference acceleration)
      real s = J[0] * x[0] + J[1] * x[1] + J[2] * x[2] + J[3] * x[3] + J[4] * x[4] + J[5] * x[5];
      if (body >= 2) { const real* xl = x + 6 + 3 * info_leg(info); s += J[6] * xl[0] + J[7] * xl[1] + J[8] * xl[2]; }
      if (minus_ar) s -= w.c_ar[c][k];
      out_c[c][k] = s;
    }
    syncwarp();
  }

  // constraint cost at the current residuals (no side effects); warp-uniform result
  QS_DEV real units_cost() {
    real cost = 0, d1 = 0, d2 = 0;
    if (lane < NSC) scalar_unit_eval(lane, w.u_r[lane], real(0), cost, d1, d2);
    real zero[MAXDIM];
    for (int k = 0; k < MAXDIM; k++) zero[k] = 0;
    static_assert(NCON <= 32, "one lane per contact");
    if (lane < w.ncon) contact_unit_eval(lane, w.c_r[lane], zero, cost, d1, d2);
    cost = warp_sum(cost);
    syncwarp();  // residuals are overwritten by the next units_Jx
    return cost;
  }

  // states / forces / weights at the current residuals (w.u_r, w.c_r); returns this lane's share of the constraint cost
  QS_DEV real units_update() {
    real cost = 0;
    if (lane < NSC) {
      const int u = lane;
      const real x = w.u_r[u], D = w.u_D[u];
      real F = 0, Wt = 0;
      if (D != 0) {
        if (u < NFL) {
          const real f = m.dof_floss[6 + u], rf = N::div(f, D);
          if (x <= -rf) { F = f; cost = -f * (real(0.5) * rf + x); }
          else if (x >= rf) { F = -f; cost = -f * (real(0.5) * rf - x); }
          else { F = -D * x; Wt = D; cost = real(0.5) * D * x * x; }
        } else if (x < 0) { F = -D * x; Wt = D; cost = real(0.5) * D * x * x; }
      }
      w.u_F[u] = F; w.u_W[u] = Wt;
    }
    const int ncon = w.ncon;
    if (lane < ncon) cost += contact_unit_update(lane);
    syncwarp();
    return cost;  // per-lane partial: the caller reduces it together with its other sums
  }

  // qfrc_constraint = J^T F (into w.fcon) and grad = Ma - fsm - fcon; dof lanes
  QS_DEV void constraint_force_and_grad() {
    if (lane < NV) {
      const int d = lane;
      const bool isleg = d >= 6;
      const int l = dof_leg(), k = dof_k(), col = isleg ? 6 + k : d, j = isleg ? d - 6 : 0;
      real s = isleg ? w.u_F[j] + (kLimits ? w.u_sign[NFL + j] * w.u_F[NFL + j] : real(0)) : real(0);
      const int ncon = w.ncon;
      for (int c = 0; c < ncon; c++) {
        const int info = w.c_info[c], body = info_body(info), dim = info_dim(info);
        const bool mine = !isleg || info_leg(info) == l;
        real t = 0;
#pragma unroll
        for (int a = 0; a < MAXDIM; a++) t += w.Jc[c][a][col] * w.c_F[c][a];  // rows >= dim: zero Jacobian, zero force
        s += mine ? t : real(0);
      }
      w.fcon[d] = s;
      w.grad[d] = w.Ma[d] - w.fsm[d] - s;
    }
    syncwarp();
  }

  // H = M + J^T W J in block form.  Only the entries the factorisation reads are produced (lower triangle of the base
  // block, the full leg-base couplings, lower triangle of the leg blocks).  Contacts are accumulated one after the other;
  // lane e owns entry (i,j), i>=j, of the contact-local symmetric 9x9 matrix Jc^T W Jc (45 entries -> two passes), so the
  // same lane always touches the same shared-memory word and no synchronisation is needed between contacts.
  QS_DEV void build_hessian() {
    const int i0 = tri & 15, j0 = (tri >> 4) & 15, i1 = (tri >> 8) & 15, j1 = (tri >> 12) & 15;
    const bool own1 = lane < 13;  // e1 = lane + 32 < 45
    real ab = 0, a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};  // base-block entry; per-leg entries of e0 (>= 21) and e1
    const int ncon = w.ncon;
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {
      const real* Wt = w.c_W[c];
      if (Wt[0] == 0) continue;  // no active row in this contact (W00 sums the active regularisers)
      const int info = w.c_info[c], body = info_body(info), dim = info_dim(info);
      const bool legc = body >= 2;
      const int leg = legc ? info_leg(info) : 0;
      real h0 = 0, h1 = 0;
      const bool on0 = lane < 21 || legc, on1 = own1 && legc;
      if (MAXDIM == 3 && cone_is_pyramidal() && dim == 3) {
        const real w00 = Wt[widx(0, 0)], w01 = Wt[widx(0, 1)], w02 = Wt[widx(0, 2)], w11 = Wt[widx(1, 1)], w22 = Wt[widx(2, 2)];
        if (on0) {
          const real p0 = w.Jc[c][0][i0], p1 = w.Jc[c][1][i0], p2 = w.Jc[c][2][i0], q0 = w.Jc[c][0][j0], q1 = w.Jc[c][1][j0], q2 = w.Jc[c][2][j0];
          h0 = w00 * p0 * q0 + w01 * (p0 * q1 + p1 * q0) + w02 * (p0 * q2 + p2 * q0) + w11 * p1 * q1 + w22 * p2 * q2;
        }
        if (on1) {
          const real p0 = w.Jc[c][0][i1], p1 = w.Jc[c][1][i1], p2 = w.Jc[c][2][i1], q0 = w.Jc[c][0][j1], q1 = w.Jc[c][1][j1], q2 = w.Jc[c][2][j1];
          h1 = w00 * p0 * q0 + w01 * (p0 * q1 + p1 * q0) + w02 * (p0 * q2 + p2 * q0) + w11 * p1 * q1 + w22 * p2 * q2;
        }
      } else if (dim == 3) {
        // elliptic cone, three rows: the full symmetric 3x3 weight written out (same summation order as the loop below)
        const real w00 = Wt[widx(0, 0)], w01 = Wt[widx(0, 1)], w02 = Wt[widx(0, 2)], w11 = Wt[widx(1, 1)], w12 = Wt[widx(1, 2)], w22 = Wt[widx(2, 2)];
        if (on0) {
          const real p0 = w.Jc[c][0][i0], p1 = w.Jc[c][1][i0], p2 = w.Jc[c][2][i0], q0 = w.Jc[c][0][j0], q1 = w.Jc[c][1][j0], q2 = w.Jc[c][2][j0];
          h0 = p0 * (w00 * q0 + w01 * q1 + w02 * q2) + p1 * (w01 * q0 + w11 * q1 + w12 * q2) + p2 * (w02 * q0 + w12 * q1 + w22 * q2);
        }
        if (on1) {
          const real p0 = w.Jc[c][0][i1], p1 = w.Jc[c][1][i1], p2 = w.Jc[c][2][i1], q0 = w.Jc[c][0][j1], q1 = w.Jc[c][1][j1], q2 = w.Jc[c][2][j1];
          h1 = p0 * (w00 * q0 + w01 * q1 + w02 * q2) + p1 * (w01 * q0 + w11 * q1 + w12 * q2) + p2 * (w02 * q0 + w12 * q1 + w22 * q2);
        }
      } else {
        // condim 1 / 6: a real loop (not unrolled): this body runs once per contact per Newton iteration in every lane, and the
        // elliptic kernels are instruction-fetch bound (no_instruction 4.6 cycles per issue, profiles/r02_cfg4_*)
#pragma unroll 1
        for (int a = 0; a < dim; a++) {
          real t0 = 0, t1 = 0;
#pragma unroll 1
          for (int b = 0; b < dim; b++) {
            const real wab = Wt[widx(a, b)];
            if (on0) t0 += wab * w.Jc[c][b][j0];
            if (on1) t1 += wab * w.Jc[c][b][j1];
          }
          if (on0) h0 += w.Jc[c][a][i0] * t0;
          if (on1) h1 += w.Jc[c][a][i1] * t1;
        }
      }
      if (lane < 21) ab += h0;
      else {
#pragma unroll
        for (int l = 0; l < 4; l++) a0[l] += (l == leg) ? h0 : real(0);
      }
#pragma unroll
      for (int l = 0; l < 4; l++) a1[l] += (l == leg) ? h1 : real(0);
    }
    // H = M + accumulated J^T W J (+ the scalar units' weights on the leg diagonals), written once by the owning lanes
    if (lane < 21) w.hes.Hbb[i0][j0] = w.Mbb[i0][j0] + ab;
    // per-leg entries: the (row, column) of a lane is fixed, only the leg stride moves (Hlb / Mlb: 18 floats per leg, Hll / Mll: 9)
    auto write_legs = [&](int i, int j, const real* acc) {
      const int k = i - 6;
      if (j < 6) {
        real* H = &w.hes.Hlb[0][k][j];
        const real* M = &w.Mlb[0][k][j];
#pragma unroll
        for (int l = 0; l < 4; l++) H[18 * l] = M[18 * l] + acc[l];
      } else {
        const int jj = j - 6;
        real* H = &w.hes.Hll[0][k][jj];
        const real* M = &w.Mll[0][k][jj];
        const bool diag = jj == k;
        const real* uw = &w.u_W[k];
#pragma unroll
        for (int l = 0; l < 4; l++) H[9 * l] = M[9 * l] + acc[l] + (diag ? uw[3 * l] + (kLimits ? uw[NFL + 3 * l] : real(0)) : real(0));
      }
    };
    if (lane >= 21) write_legs(i0, j0, a0);
    if (own1) write_legs(i1, j1, a1);
    syncwarp();
  }

  struct LsPoint { real alpha, cost, d1, d2; };

  // cost and derivatives along qacc + alpha*search (warp-uniform result)
  QS_DEV LsPoint ls_eval(real alpha, real qg0, real qg1, real qg2) {
    real cost = 0, d1 = 0, d2 = 0;
    if (lane < NSC) scalar_unit_eval(lane, w.u_r[lane] + alpha * w.u_v[lane], w.u_v[lane], cost, d1, d2);
    const int ncon = w.ncon;
    if (lane < ncon) {
      const int c = lane;
      real r[MAXDIM];
      for (int k = 0; k < MAXDIM; k++) r[k] = w.c_r[c][k] + alpha * w.c_v[c][k];
      contact_unit_eval(c, r, w.c_v[c], cost, d1, d2);
    }
    LsPoint p;
    p.alpha = alpha;
    warp_sum3(cost, d1, d2);
    p.cost = cost + alpha * alpha * qg2 + alpha * qg1 + qg0;
    p.d1 = d1 + 2 * alpha * qg2 + qg1;
    p.d2 = d2 + 2 * qg2;
    return p;
  }

  // [MJ] PrimalSearch: exact line search by safeguarded 1-D Newton with bracketing.  Written as a state machine around a
  // single evaluation site to keep the code footprint (and I-cache pressure) small.
  //   state 0: value at 0; 1: first Newton point; 2: one-sided Newton until the derivative changes sign; 3: bracketed
  //   refinement with Newton steps from both ends plus the midpoint.
  QS_DEV real line_search(real gtol, real qg0, real qg1, real qg2, real cost0, real slope0) {
    constexpr real kNoise = sizeof(real) == 4 ? real(1e-6) : real(1e-14);
    int state = 1, it = 0, dir = 1, ci = 3;
    LsPoint p0{}, p1{}, p2{}, lo{}, hi{};
    // The point alpha = 0 needs no evaluation: its cost is the current cost, its slope is grad . search and, because the search
    // direction solves H s = -grad with the exact Hessian of the piecewise-quadratic cost, its curvature is s^T H s = -slope,
    // so the first Newton point is alpha = 1.
    p0.alpha = 0; p0.cost = cost0; p0.d1 = slope0; p0.d2 = -slope0;
    if (!(p0.d2 > 0)) return 0;
    // the slope cannot be resolved below a few ulps of its initial magnitude in this precision
    gtol = N::max(gtol, kNoise * N::abs(slope0));
    real a = 1;
    real cand[3] = {0, 0, 0};
    bool moved = true;
    const int ls_iter = m.ls_iterations;
    for (;;) {
      const LsPoint p = ls_eval(a, qg0, qg1, qg2);
      ls_evals++;
      if (state == 1) {
        p1 = (p0.cost < p.cost) ? p0 : p;
        if (N::abs(p1.d1) < gtol) return p1.alpha;
        dir = p1.d1 < 0 ? 1 : -1;
        p2 = p1;
        state = 2;
      } else if (state == 2) {
        p2 = p1; p1 = p; it++;
        if (N::abs(p1.d1) < gtol) return p1.alpha;
      } else {
        it++;
        if (N::abs(p.d1) < gtol) return p.alpha;
        if ((p.d1 < 0) == (lo.d1 < 0)) lo = p; else hi = p;
        moved = true;
      }
      if (state == 2) {
        if (p1.d1 * dir <= -gtol && it < ls_iter) { a = p1.alpha - N::div(p1.d1, p1.d2); continue; }
        if (it >= ls_iter) return p1.alpha;
        lo = p2; hi = p1; state = 3; ci = 3; moved = true;
      }
      bool have = false;
      while (!have) {
        if (ci >= 3) {
          if (!moved || it >= ls_iter) return lo.cost < hi.cost ? lo.alpha : hi.alpha;
          cand[0] = lo.alpha - N::div(lo.d1, lo.d2); cand[1] = hi.alpha - N::div(hi.d1, hi.d2); cand[2] = real(0.5) * (lo.alpha + hi.alpha);
          ci = 0;
          moved = false;
        }
        const real c = cand[ci++];
        if ((c > lo.alpha && c < hi.alpha) || (c < lo.alpha && c > hi.alpha)) { a = c; have = true; }
      }
    }
  }

  // [MJ] mj_fwdConstraint + mj_solNewton
  QS_DEV void solve(int max_iter, real tol) {
    QS_T0();
    // qacc_smooth = M^-1 qfrc_smooth
    copy_M_to_H(real(0));
    factor_H(w, lane, tri);
    solve_H(w, lane, tri, w.fsm, w.asmooth, real(1));
    // warm start: cheaper of qacc_warmstart and qacc_smooth
    units_Jx(w.warm, w.u_r, w.c_r, true);
    real cost_w = units_cost();
    {
      const real ma = mass_matvec(w, w.warm, lane, tri);
      cost_w += warp_sum((lane < NV) ? real(0.5) * (ma - w.fsm[lane]) * (w.warm[lane] - w.asmooth[lane]) : real(0));
    }
    units_Jx(w.asmooth, w.u_r, w.c_r, true);
    const real cost_s = units_cost();
    const bool use_warm = cost_w < cost_s;
    if (lane < NV) w.qacc[lane] = use_warm ? w.warm[lane] : w.asmooth[lane];
    syncwarp();
    if (use_warm) units_Jx(w.qacc, w.u_r, w.c_r, true);
    {
      const real ma = mass_matvec(w, w.qacc, lane, tri);
      if (lane < NV) w.Ma[lane] = ma;
    }
    syncwarp();
    const real scale = real(1) / (m.meaninertia * NV);
    int iter = 0;
    real cost = 0, oldcost = 0, pred_prev = 0;
    solver_maxed = false;
    QS_TACC(0);
    while (true) {
      // cost, forces, gradient at the current point
      real ccost = units_update();
      const int dl_ = lane < NV ? lane : 0;
      real gauss = (lane < NV) ? real(0.5) * (w.Ma[dl_] - w.fsm[dl_]) * (w.qacc[dl_] - w.asmooth[dl_]) : real(0);
      warp_sum2(ccost, gauss);
      oldcost = cost;
      cost = gauss + ccost;
      constraint_force_and_grad();
      if (iter > 0) {
        // [MJ] stop on small scaled improvement or gradient; in addition stop when either is below what this precision can
        // resolve (a few ulps of the cost / of the force balance) -- otherwise fp32 rounding noise of random sign keeps the
        // loop alive with ~50% probability per iteration.
        constexpr real kNoise = sizeof(real) == 4 ? real(1e-6) : real(1e-14);    // ~16 ulp: force-balance residual floor
        constexpr real kNoiseCost = sizeof(real) == 4 ? real(1e-7) : real(1e-15);  // ~2 ulp: cost did not move at all
        real gn2 = (lane < NV) ? w.grad[dl_] * w.grad[dl_] : real(0);
        real fn2 = (lane < NV) ? w.Ma[dl_] * w.Ma[dl_] + w.fsm[dl_] * w.fsm[dl_] + w.fcon[dl_] * w.fcon[dl_] : real(0);
        warp_sum2(gn2, fn2);
        const real imp = oldcost - cost;
#ifdef QS_PROF_IMP
        if (iter <= 8) tacc[iter - 1] = __float_as_uint(float(scale * imp));
#endif
        // The measured improvement of a stiff problem (tiny friction -> huge pyramid regularisers) is a difference of two large
        // costs: once it drops below what this precision resolves it reads as "no progress" long before the solution is reached.
        // In that regime the decrease predicted by the Newton step just taken (-1/2 grad . search: built from the gradient, not
        // from cost differences) stands in for it.
        const bool resolved = N::abs(imp) > kNoiseCost * (N::abs(cost) + N::abs(oldcost));
        const real imp_eff = (resolved || sizeof(real) == 8) ? imp : pred_prev;
        if (scale * imp_eff < tol || (sizeof(real) == 8 && !resolved)) break;
        if (scale * scale * gn2 < tol * tol || gn2 < kNoise * kNoise * fn2) break;
      }
      if (iter >= max_iter) { solver_maxed = true; break; }
      QS_TACC(1);
      // Newton direction
#ifdef QS_PROF_REPEAT
#pragma unroll 1
      for (int rep = 0; rep < 2; rep++) {
        build_hessian();
        factor_H(w, lane, tri);
        solve_H(w, lane, tri, w.grad, w.search, real(-1));
        if (rep == 0) QS_TACC(2); else QS_TACC(3);
      }
#else
      build_hessian();
      QS_TACC(2);
      factor_H(w, lane, tri);
      QS_TACC(3);
      solve_H(w, lane, tri, w.grad, w.search, real(-1));
      QS_TACC(4);
#endif
      // exact line search
      const real mv = mass_matvec(w, w.search, lane, tri);
      if (lane < NV) w.Mv[lane] = mv;
      units_Jx(w.search, w.u_v, w.c_v, false);
      const real sd = (lane < NV) ? w.search[dl_] : real(0);
      real sn2 = sd * sd, qg1 = sd * (w.Ma[dl_] - w.fsm[dl_]), qg2 = real(0.5) * sd * mv, slope0 = w.grad[dl_] * sd;
      warp_sum4(sn2, qg1, qg2, slope0);
      const real snorm = N::sqrt(sn2);
      if (snorm < N::minval) break;
      pred_prev = real(-0.5) * slope0;  // Newton decrement of this step
      const real gtol = tol * m.ls_tolerance * snorm / scale;
      QS_TACC(5);
      const real alpha = line_search(gtol, gauss, qg1, qg2, cost, slope0);
      QS_TACC(6);
      syncwarp();  // the last evaluation read the residuals that the move below updates
      if (alpha == 0) break;
      // move
      if (lane < NV) { w.qacc[lane] += alpha * w.search[lane]; w.Ma[lane] += alpha * w.Mv[lane]; }
      if (lane < NSC) w.u_r[lane] += alpha * w.u_v[lane];
      for (int it2 = lane; it2 < w.ncon * MAXDIM; it2 += 32) (&w.c_r[0][0])[it2] += alpha * (&w.c_v[0][0])[it2];
      syncwarp();
      iter++;
      QS_TACC(7);
    }
    QS_TACC(1);
    solver_iter = iter;
  }

  // accelerometer / gyro at the IMU site. [MJ] mj_rnePostConstraint + mj_objectAcceleration
  QS_DEV void sensors() {
    if (lane == 0) {
      const real* st = w.sens_tmp;
      real cv[6], ca[6];
      for (int i = 0; i < 6; i++) { cv[i] = st[i]; ca[i] = st[6 + i]; }
      for (int d = 0; d < 6; d++) for (int i = 0; i < 6; i++) ca[i] += w.cdof[d][i] * w.qacc[d];
      const real* spos = st + 12;
      const real* R = st + 15;
      real vl[3], al[3], c1[3];
      cross3(c1, cv, spos);
      for (int i = 0; i < 3; i++) vl[i] = cv[3 + i] + c1[i];
      cross3(c1, ca, spos);
      for (int i = 0; i < 3; i++) al[i] = ca[3 + i] + c1[i];
      cross3(c1, cv, vl);
      for (int i = 0; i < 3; i++) al[i] += c1[i];
      mul_mtv(w.sens, R, al);
      mul_mtv(w.sens + 3, R, cv);
    }
    syncwarp();
  }

  // position-dependent half of the forward pass (body poses stay valid in w.kin until forward_dynamics starts the solver)
  QS_DEV void forward_position() {
    kinematics();
    com_inertia();
    cdof();
    collide_floor();
  }
  // velocity / force half: bias, mass matrix, constraints, solver
  QS_DEV void forward_dynamics(int max_iter, real tol) {
    bias_and_smooth();
    mass_matrix();
    make_constraints();
    solve(max_iter, tol);
    if (has_imu()) sensors();
  }
  QS_DEV void forward(int max_iter, real tol) {
    forward_position();
    forward_dynamics(max_iter, tol);
  }

  // [MJ] mj_Euler with implicit joint damping. org = internal-frame origin for the fp64 base position.
  QS_DEV void integrate(double* base64) {
    const real h = m.timestep;
    copy_M_to_H(h);
    factor_H(w, lane, tri);
    if (lane < NV) w.grad[lane] = w.fsm[lane] + w.fcon[lane];
    syncwarp();
    solve_H(w, lane, tri, w.grad, w.search, real(1));
    if (lane < NV) w.qvel[lane] += h * w.search[lane];
    syncwarp();
    if (lane < 3) {
      const double p = (lane < 2 ? w.org[lane] : 0.0) + double(w.qpos[lane]) + double(h) * double(w.qvel[lane]);
      base64[lane] = p;
      w.qpos[lane] = real(p - (lane < 2 ? w.org[lane] : 0.0));
    } else if (lane == 3) {
      const real wx = w.qvel[3], wy = w.qvel[4], wz = w.qvel[5];
      const real nrm = N::sqrt(wx * wx + wy * wy + wz * wz), ang = h * nrm;
      real q[4] = {w.qpos[3], w.qpos[4], w.qpos[5], w.qpos[6]};
      if (ang > 0) {
        real s, c;
        N::sincos(real(0.5) * ang, &s, &c);
        const real qr[4] = {c, s * wx / nrm, s * wy / nrm, s * wz / nrm};
        real q2[4];
        quat_mul(q2, q, qr);
        for (int i = 0; i < 4; i++) q[i] = q2[i];
      }
      quat_normalize(q);
      for (int i = 0; i < 4; i++) w.qpos[3 + i] = q[i];
    } else if (lane >= 7 && lane < NQ) {
      w.qpos[lane] += h * w.qvel[lane - 1];
    }
    syncwarp();
  }

  // ------------------------------------------------------------------ env side: flags + ALL_OBS (quadruped_env.py:1146-1257)
  struct Flags { unsigned contact_mask; unsigned invalid_mask; bool out_of_bounds; };
  QS_DEV Flags flags() const {
    unsigned cm = cm_acc, im = im_acc;
    for (int o = 16; o > 0; o >>= 1) { cm |= shfl_xor(cm, o); im |= shfl_xor(im, o); }
    Flags f;
    f.contact_mask = cm; f.invalid_mask = im;
    f.out_of_bounds = out_of_bounds();
    return f;
  }
  // _check_out_of_terrain_bounds (quadruped_env.py:1250-1257) on the current base position
  QS_DEV bool out_of_bounds() const {
    const double x = w.org[0] + double(w.qpos[0]), y = w.org[1] + double(w.qpos[1]);
    return x > double(m.terrain_limits[0]) || x < double(m.terrain_limits[1]) || y > double(m.terrain_limits[2]) || y < double(m.terrain_limits[3]);
  }

  // Packs the 227 ALL_OBS scalars into w.obs (layout: SURVEY.md section 8a). Mixed time levels as in the reference
  // (App. B.1): qpos/qvel are post-integration, foot positions / Jacobians / contacts / qacc / M are from the forward pass.
  // w.obs recycles the contact-Jacobian storage, so this must be the last consumer of the solver state.
  QS_DEV void pack_obs(const real* command, unsigned contact_mask) {
    real q[4] = {w.qpos[3], w.qpos[4], w.qpos[5], w.qpos[6]}, R[9];
    quat_normalize(q);
    quat_to_mat(R, q);
    const real yaw = N::atan2(R[3], R[0]);
    // cos / sin of yaw = atan2(R10, R00) without trigonometry (heading frame of :990-997)
    const real h2 = R[0] * R[0] + R[3] * R[3];
    const real hinv = h2 > N::minval ? N::rsqrt(h2) : real(0);
    const real cy = h2 > N::minval ? R[0] * hinv : real(1), sy = R[3] * hinv;
    const real vref[3] = {cy * command[0] - sy * command[1], sy * command[0] + cy * command[1], command[2]};
    const real* v = w.qvel;
    const real* wb = w.qvel + 3;
    real* o = w.obs;
    const real ox = real(w.org[0]), oy = real(w.org[1]);
    // kinetic energy / work first: they need nothing from the recycled region
    const real mvv = mass_matvec(w, w.qvel, lane, tri), maa = mass_matvec(w, w.qacc, lane, tri);
    const int dl_ = lane < NV ? lane : 0;
    real ke = (lane < NV) ? real(0.5) * w.qvel[dl_] * mvv : real(0), wk = (lane < NV) ? maa * w.qvel[dl_] : real(0);
    warp_sum2(ke, wk);
    // spatial velocity of each calf body with the post-step qvel (feet_vel = J qvel, :652-664): lanes (leg, component), parked in the
    // IMU scratch (consumed by sensors() earlier in the step)
    if (lane < 24) {
      const int l = (tri >> 20) & 7, i = (tri >> 23) & 7;
      real sv = 0;
      for (int d = 0; d < 6; d++) sv += w.cdof[d][i] * w.qvel[d];
      for (int k = 0; k < 3; k++) { const int d = 6 + 3 * l + k; sv += w.cdof[d][i] * w.qvel[d]; }
      w.sens_tmp[lane] = sv;
    }
    syncwarp();
    // base block: lane i < 3 owns component i of every 3-vector (rows / columns of R picked without indexing the register array)
    if (lane < 3) {
      const int i = lane;
      const real Ri0 = i == 0 ? R[0] : (i == 1 ? R[3] : R[6]), Ri1 = i == 0 ? R[1] : (i == 1 ? R[4] : R[7]), Ri2 = i == 0 ? R[2] : (i == 1 ? R[5] : R[8]);
      const real Ci0 = i == 0 ? R[0] : (i == 1 ? R[1] : R[2]), Ci1 = i == 0 ? R[3] : (i == 1 ? R[4] : R[5]), Ci2 = i == 0 ? R[6] : (i == 1 ? R[7] : R[8]);
      const real vi = v[i], wbi = wb[i], vrefi = i == 0 ? vref[0] : (i == 1 ? vref[1] : vref[2]), wrefi = i == 2 ? command[3] : real(0);
      const real Rwb = Ri0 * wb[0] + Ri1 * wb[1] + Ri2 * wb[2];
      o[i] = i < 2 ? real(w.org[i] + double(w.qpos[i])) : w.qpos[2];
      o[3 + i] = vi; o[6 + i] = vrefi - vi; o[9 + i] = w.qacc[i];
      o[12 + i] = Rwb; o[15 + i] = wrefi - Rwb;
      o[18 + i] = i == 0 ? N::atan2(R[7], R[8]) : (i == 1 ? -N::asin(N::max(real(-1), N::min(real(1), R[6]))) : yaw);
      o[25 + 3 * i] = Ri0; o[26 + 3 * i] = Ri1; o[27 + 3 * i] = Ri2;
      o[34 + i] = -Ci2;
      const real RTv = Ci0 * v[0] + Ci1 * v[1] + Ci2 * v[2], RTvref = Ci0 * vref[0] + Ci1 * vref[1] + Ci2 * vref[2];
      o[37 + i] = RTv; o[40 + i] = RTvref - RTv;
      o[43 + i] = Ci0 * w.qacc[0] + Ci1 * w.qacc[1] + Ci2 * w.qacc[2];
      o[46 + i] = wbi;
      o[49 + i] = (Ci0 * real(0) + Ci1 * real(0) + Ci2 * command[3]) - wbi;
    } else if (lane == 3) {
      for (int i = 0; i < 4; i++) o[21 + i] = w.qpos[3 + i];
      o[125] = ke; o[126] = wk;
    }
    if (lane < NQ) o[52 + lane] = (lane < 2) ? real(w.org[lane] + double(w.qpos[lane])) : w.qpos[lane];
    if (lane < NV) o[71 + lane] = w.qvel[lane];
    if (lane < NU) { o[89 + lane] = w.ctrl[lane]; o[101 + lane] = w.qpos[7 + lane]; o[113 + lane] = w.qvel[6 + lane]; }
    // feet block: lane (leg l, component c) on the dof lanes 6..17
    if (lane >= 6 && lane < NV) {
      const int l = dof_leg(), c = dof_k();
      const real* cv6 = w.sens_tmp + 6 * l;
      const real* p = w.footpos[l];
      const real off[3] = {p[0] - w.com[0], p[1] - w.com[1], p[2] - w.com[2]};
      real cr[3], fv[3], fr[3];
      cross3(cr, cv6, off);
      for (int i = 0; i < 3; i++) fv[i] = cv6[3 + i] + cr[i];
      const real rb[3] = {p[0] - w.qpos[0], p[1] - w.qpos[1], p[2] - w.qpos[2]};
      cross3(cr, wb, rb);
      for (int i = 0; i < 3; i++) fr[i] = fv[i] - v[i] - cr[i];
      const real Cc0 = c == 0 ? R[0] : (c == 1 ? R[1] : R[2]), Cc1 = c == 0 ? R[3] : (c == 1 ? R[4] : R[5]), Cc2 = c == 0 ? R[6] : (c == 1 ? R[7] : R[8]);
      const int k = 3 * l + c;
      o[127 + k] = p[c] + (c == 0 ? ox : (c == 1 ? oy : real(0)));
      o[139 + k] = Cc0 * rb[0] + Cc1 * rb[1] + Cc2 * rb[2];
      o[151 + k] = c == 0 ? fv[0] : (c == 1 ? fv[1] : fv[2]);
      o[163 + k] = c == 0 ? fr[0] : (c == 1 ? fr[1] : fr[2]);
      o[175 + k] = Cc0 * fv[0] + Cc1 * fv[1] + Cc2 * fv[2];
      o[187 + k] = Cc0 * fr[0] + Cc1 * fr[1] + Cc2 * fr[2];
      if (c == 0) o[199 + l] = (contact_mask >> l) & 1u ? real(1) : real(0);
    } else if (lane >= 18 && lane < 22) {
      const int l = lane - 18;
      real cf[3] = {0, 0, 0}, t[3];
      for (int c = 0; c < w.ncon; c++) {
        const int info = w.c_info[c];
        if (info_body(info) != 4 + 3 * l) continue;
        const int dim = info_dim(info);
        const real F0 = w.c_F[c][0], F1 = dim > 1 ? w.c_F[c][1] : real(0), F2 = dim > 1 ? w.c_F[c][2] : real(0);
        real t2[3];
        cross3(t2, w.c_frame[c], w.c_frame[c] + 3);
        for (int i = 0; i < 3; i++) cf[i] += w.c_frame[c][i] * F0 + w.c_frame[c][3 + i] * F1 + t2[i] * F2;
      }
      mul_mtv(t, R, cf);
      for (int i = 0; i < 3; i++) { o[203 + 3 * l + i] = cf[i]; o[215 + 3 * l + i] = t[i]; }
    }
    syncwarp();
  }
};

}  // namespace qs
