/*
 * qstep_oracle.c -- TEST INFRASTRUCTURE ONLY. fp64, scalar, single-env CPU restatement of the hot path of
 * iit-DLSLab/gym-quadruped: `mujoco.mj_step` as called at gym_quadruped/quadruped_env.py:271,397 followed by the
 * env-side observation pack (`_get_obs`, quadruped_env.py:1146-1226), termination checks (:1228-1257), the reset lift
 * loop (:376-388), IMU truth signals (sensors/imu.py:110-139) and the height-map ray cast (sensors/heightmap.py:66-104).
 *
 * Third-party attribution: the formulas marked [MJ] restate published algorithms of the MuJoCo physics engine (Google DeepMind,
 * Apache License 2.0), the engine the reference calls through its `mujoco` dependency.  MuJoCo's sources are not part of
 * /root/reference and nothing was copied from them; the restatements follow the engine's documentation and SURVEY.md App. A.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party engine MuJoCo (pyproject.toml:30,
 * `mujoco>=3.10.0`, no lock file), which is neither vendored under /root/reference nor installable in this image, and
 * the reference's only test (tests/env_test.py:14-53) holds no golden numbers.  This file restates the engine's
 * published pipeline (every engine-specific formula is marked [MJ]) and is pinned only by
 * analytic known-answers (tests/test_oracle_physics.py) and by the reference's own Python-side formulas.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (gym_quadruped_b200/csrc) never links or calls it.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/qstep.h"

#define NB QS_NBODY
#define NV QS_NV
#define NQ QS_NQ
#define NU QS_NU
#define MAXCON 208 /* more than the collision stage can produce (48 geoms x 4 + slack): the oracle never truncates its contact set */
#define MAXEFC (NV + 2 * QS_NJNT + MAXCON * 10)
#define MINVAL 1e-15 /* [MJ] mjMINVAL */
#define MINMU 1e-5   /* [MJ] mjMINMU  */
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { T_FRICTION = 0, T_LIMIT = 1, T_CONTACT_FRICTIONLESS = 2, T_CONTACT_PYRAMIDAL = 3, T_CONTACT_ELLIPTIC = 4 };
enum { S_SATISFIED = 0, S_QUADRATIC = 1, S_LINEARNEG = 2, S_LINEARPOS = 3, S_CONE = 4 };

typedef struct {
  double dist, pos[3], frame[9];
  int geom;  /* robot geom index                                            */
  int body;  /* robot body                                                  */
  int wgeom; /* world geom: 0 = floor plane, 1 = hfield, 1.. = boxes        */
  double sign; /* +1: world geom is geom1 (normal world->robot); -1: robot geom is geom1 */
  double friction[5], solref[2], solimp[5], includemargin, mu;
  int dim, efc_address, exclude;
  double force[6]; /* contact-frame wrench, as mj_contactForce (quadruped_env.py:852) */
} OContact;

typedef struct {
  QsModel m;
  double* vert;
  float* hf;
  /* state */
  double qpos[NQ], qvel[NV], ctrl[NU], qfrc_applied[NV], qacc_warmstart[NV], time;
  double mu_floor, mu_feet; /* <0: model values (quadruped_env.py:1277-1296 not yet called) */
  double command[4];
  /* position stage */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  double xanchor[QS_NJNT][3], xaxis[QS_NJNT][3];
  double geom_xpos[QS_MAXGEOM][3], geom_xmat[QS_MAXGEOM][9];
  double com[3];
  double cinert[NB][10], crb[NB][10], cdof[NV][6], cdof_dot[NV][6], cvel[NB][6], cacc[NB][6], cfrc[NB][6];
  double M[NV][NV], L[NV][NV];
  double qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV], qacc[NV],
      qfrc_constraint[NV];
  int ncon;
  OContact con[MAXCON];
  int nefc;
  int efc_type[MAXEFC], efc_id[MAXEFC], efc_state[MAXEFC];
  double efc_J[MAXEFC][NV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_floss[MAXEFC], efc_diagApprox[MAXEFC],
      efc_R[MAXEFC], efc_D[MAXEFC], efc_vel[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC], efc_imp[MAXEFC];
  int solver_iter, overflow;
  void* ctx; /* solver scratch (Ctx), per handle so handles are thread-independent */
  double sensor_acc[3], sensor_gyro[3];
  /* info */
  int contact_state[4], invalid_contact, out_of_bounds;
  unsigned invalid_body_mask;
} OData;

/* ------------------------------------------------------------------ small math */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static void mulMatVec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulMatTVec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void quatMul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quatNormalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat2Mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void rotVecQuat(double* r, const double* v, const double* q) {
  double m[9];
  quat2Mat(m, q);
  mulMatVec3(r, m, v);
}
/* [MJ] mju_makeFrame: frame[0:3] = normal given, frame[3:6] optional guess */
static void makeFrame(double* f) {
  double n = norm3(f);
  for (int i = 0; i < 3; i++) f[i] /= n;
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double d = dot3(f, f + 3);
  for (int i = 0; i < 3; i++) f[3 + i] -= d * f[i];
  n = norm3(f + 3);
  for (int i = 0; i < 3; i++) f[3 + i] /= n;
  cross3(f + 6, f, f + 3);
}
/* spatial algebra in the engine's [rot(3), lin(3)] convention, com-based frame */
static void mulInertVec(double* r, const double* I, const double* v) {
  r[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] - I[8] * v[4] + I[7] * v[5];
  r[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + I[8] * v[3] - I[6] * v[5];
  r[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] - I[7] * v[3] + I[6] * v[4];
  r[3] = I[8] * v[1] - I[7] * v[2] + I[9] * v[3];
  r[4] = I[6] * v[2] - I[8] * v[0] + I[9] * v[4];
  r[5] = I[7] * v[0] - I[6] * v[1] + I[9] * v[5];
}
static void crossMotion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v);
  cross3(b, vel, v + 3);
  cross3(c, vel + 3, v);
  for (int i = 0; i < 3; i++) { r[i] = a[i]; r[3 + i] = b[i] + c[i]; }
}
static void crossForce(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  cross3(c, vel, f + 3);
  for (int i = 0; i < 3; i++) { r[i] = a[i] + b[i]; r[3 + i] = c[i]; }
}

/* ------------------------------------------------------------------ position stage */
static int dof_body(int d) { return d < 6 ? 1 : d - 6 + 2; }

/* [MJ] mj_kinematics, SURVEY App. A.1 */
static void kinematics(OData* d) {
  const QsModel* m = &d->m;
  memset(d->xpos[0], 0, sizeof(d->xpos[0]));
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2Mat(d->xmat[0], d->xquat[0]);
  for (int b = 1; b < NB; b++) {
    int p = m->body_parent[b];
    double pos[3], quat[4], tmp[3];
    if (b == 1) { /* free joint: pose straight from qpos, quaternion normalised */
      memcpy(pos, d->qpos, sizeof(pos));
      memcpy(quat, d->qpos + 3, sizeof(quat));
      quatNormalize(quat);
    } else {
      int j = b - 2;
      mulMatVec3(tmp, d->xmat[p], m->body_pos[b]);
      for (int i = 0; i < 3; i++) pos[i] = d->xpos[p][i] + tmp[i];
      quatMul(quat, d->xquat[p], m->body_quat[b]);
      /* hinge: anchor & axis in the frame before the joint rotation; angle = qpos - qpos0 */
      rotVecQuat(tmp, m->jnt_pos[j], quat);
      for (int i = 0; i < 3; i++) d->xanchor[j][i] = pos[i] + tmp[i];
      rotVecQuat(d->xaxis[j], m->jnt_axis[j], quat);
      double ang = d->qpos[7 + j] - m->qpos0[7 + j];
      double s = sin(0.5 * ang), qloc[4] = {cos(0.5 * ang), s * m->jnt_axis[j][0], s * m->jnt_axis[j][1], s * m->jnt_axis[j][2]};
      double q2[4];
      quatMul(q2, quat, qloc);
      memcpy(quat, q2, sizeof(quat));
      rotVecQuat(tmp, m->jnt_pos[j], quat); /* off-centre correction */
      for (int i = 0; i < 3; i++) pos[i] = d->xanchor[j][i] - tmp[i];
    }
    quatNormalize(quat);
    memcpy(d->xpos[b], pos, sizeof(pos));
    memcpy(d->xquat[b], quat, sizeof(quat));
    quat2Mat(d->xmat[b], quat);
    mulMatVec3(tmp, d->xmat[b], m->body_ipos[b]);
    for (int i = 0; i < 3; i++) d->xipos[b][i] = pos[i] + tmp[i];
    double qi[4];
    quatMul(qi, quat, m->body_iquat[b]);
    quat2Mat(d->ximat[b], qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double tmp[3], q[4];
    mulMatVec3(tmp, d->xmat[b], m->geom_pos[g]);
    for (int i = 0; i < 3; i++) d->geom_xpos[g][i] = d->xpos[b][i] + tmp[i];
    quatMul(q, d->xquat[b], m->geom_quat[g]);
    quat2Mat(d->geom_xmat[g], q);
  }
}

/* [MJ] mj_comPos: subtree com of the root, cinert, cdof */
static void comPos(OData* d) {
  const QsModel* m = &d->m;
  double mass = 0, c[3] = {0, 0, 0};
  for (int b = 1; b < NB; b++) {
    mass += m->body_mass[b];
    for (int i = 0; i < 3; i++) c[i] += m->body_mass[b] * d->xipos[b][i];
  }
  for (int i = 0; i < 3; i++) d->com[i] = c[i] / mass;
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  for (int b = 1; b < NB; b++) {
    double o[3], *R = d->ximat[b], *I = d->cinert[b];
    const double* in = m->body_inertia[b];
    double mb = m->body_mass[b];
    for (int i = 0; i < 3; i++) o[i] = d->xipos[b][i] - d->com[i];
    /* R diag(in) R^T */
    double A[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) A[i][j] = R[3 * i] * in[0] * R[3 * j] + R[3 * i + 1] * in[1] * R[3 * j + 1] + R[3 * i + 2] * in[2] * R[3 * j + 2];
    double oo = dot3(o, o);
    I[0] = A[0][0] + mb * (oo - o[0] * o[0]);
    I[1] = A[1][1] + mb * (oo - o[1] * o[1]);
    I[2] = A[2][2] + mb * (oo - o[2] * o[2]);
    I[3] = A[0][1] - mb * o[0] * o[1];
    I[4] = A[0][2] - mb * o[0] * o[2];
    I[5] = A[1][2] - mb * o[1] * o[2];
    I[6] = mb * o[0]; I[7] = mb * o[1]; I[8] = mb * o[2]; I[9] = mb;
  }
  /* cdof: free joint = 3 world translations then 3 body-frame rotations about the base origin */
  memset(d->cdof, 0, sizeof(d->cdof));
  double off[3];
  for (int i = 0; i < 3; i++) { d->cdof[i][3 + i] = 1; off[i] = d->com[i] - d->xpos[1][i]; }
  for (int k = 0; k < 3; k++) {
    double ax[3] = {d->xmat[1][k], d->xmat[1][3 + k], d->xmat[1][6 + k]};
    memcpy(d->cdof[3 + k], ax, sizeof(ax));
    cross3(d->cdof[3 + k] + 3, ax, off);
  }
  for (int j = 0; j < QS_NJNT; j++) {
    for (int i = 0; i < 3; i++) off[i] = d->com[i] - d->xanchor[j][i];
    memcpy(d->cdof[6 + j], d->xaxis[j], 3 * sizeof(double));
    cross3(d->cdof[6 + j] + 3, d->xaxis[j], off);
  }
}

static int dof_parent(int dd) { /* [MJ] dof_parentid */
  if (dd == 0) return -1;
  if (dd < 6) return dd - 1;
  int k = (dd - 6) % 3;
  return k == 0 ? 5 : dd - 1;
}

/* [MJ] mj_crb: composite rigid body mass matrix + armature */
static void crb(OData* d) {
  const QsModel* m = &d->m;
  memcpy(d->crb, d->cinert, sizeof(d->crb));
  for (int b = NB - 1; b > 1; b--) {
    int p = m->body_parent[b];
    for (int i = 0; i < 10; i++) d->crb[p][i] += d->crb[b][i];
  }
  memset(d->M, 0, sizeof(d->M));
  for (int i = 0; i < NV; i++) {
    double buf[6];
    mulInertVec(buf, d->crb[dof_body(i)], d->cdof[i]);
    for (int j = i; j >= 0; j = dof_parent(j)) {
      double v = 0;
      for (int k = 0; k < 6; k++) v += d->cdof[j][k] * buf[k];
      d->M[i][j] = d->M[j][i] = v;
    }
    d->M[i][i] += m->dof_armature[i];
  }
}

/* dense Cholesky A = L L^T (lower); returns 0 on success */
static int cholFactor(double L[NV][NV], double A[NV][NV], int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) { if (s < MINVAL) return 1; L[i][i] = sqrt(s); }
      else L[i][j] = s / L[j][j];
    }
  return 0;
}
static void cholSolve(double L[NV][NV], double* x, const double* b, int n) {
  double y[NV];
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
}

/* ------------------------------------------------------------------ Jacobians  [MJ] mj_jac */
static void jacPoint(const OData* d, double jp[3][NV], double jr[3][NV], const double* point, int body) {
  double off[3];
  for (int i = 0; i < 3; i++) off[i] = point[i] - d->com[i];
  memset(jp, 0, 3 * NV * sizeof(double));
  if (jr) memset(jr, 0, 3 * NV * sizeof(double));
  int b = body;
  while (b > 0) {
    int d0 = (b == 1) ? 0 : 6 + (b - 2), nd = (b == 1) ? 6 : 1;
    for (int k = d0; k < d0 + nd; k++) {
      double c[3];
      cross3(c, d->cdof[k], off);
      for (int i = 0; i < 3; i++) { jp[i][k] = d->cdof[k][3 + i] + c[i]; if (jr) jr[i][k] = d->cdof[k][i]; }
    }
    b = d->m.body_parent[b];
  }
}

/* [MJ] mj_jacDot (quadruped_env.py:785): time derivative of the point Jacobian of `body` at `point`.  Needs comVel results.
 * Free-joint rotations (dofs 3..5) carry no stored cdof_dot: it is rebuilt from the full body velocity, as the engine does for
 * quaternion dofs.  Pinned by tests/test_oracle_physics.py against a finite difference of jacPoint along the flow of qvel. */
static void jacDotPoint(const OData* d, double jp[3][NV], double jr[3][NV], const double* point, int body) {
  double off[3], pvel[3], c[3];
  for (int i = 0; i < 3; i++) off[i] = point[i] - d->com[i];
  cross3(c, d->cvel[body], off);
  for (int i = 0; i < 3; i++) pvel[i] = d->cvel[body][3 + i] + c[i];
  memset(jp, 0, 3 * NV * sizeof(double));
  if (jr) memset(jr, 0, 3 * NV * sizeof(double));
  int b = body;
  while (b > 0) {
    int d0 = (b == 1) ? 0 : 6 + (b - 2), nd = (b == 1) ? 6 : 1;
    for (int k = d0; k < d0 + nd; k++) {
      double cdd[6], t1[3], t2[3];
      memcpy(cdd, d->cdof_dot[k], sizeof(cdd));
      if (b == 1 && k >= 3) crossMotion(cdd, d->cvel[1], d->cdof[k]);
      cross3(t1, cdd, off);
      cross3(t2, d->cdof[k], pvel);
      for (int i = 0; i < 3; i++) { jp[i][k] = cdd[3 + i] + t1[i] + t2[i]; if (jr) jr[i][k] = cdd[i]; }
    }
    b = d->m.body_parent[b];
  }
}

/* ------------------------------------------------------------------ collision */
static void mixParams(const QsGeomParams* a /*world*/, const QsGeomParams* b /*robot*/, const double* fa, const double* fb,
                      OContact* c) {
  /* [MJ] mj_contactParam; fa/fb = (possibly overridden) friction triples */
  double fri[3];
  if (a->priority == b->priority) {
    c->dim = a->condim > b->condim ? a->condim : b->condim;
    double mix;
    if (a->solmix >= MINVAL && b->solmix >= MINVAL) mix = a->solmix / (a->solmix + b->solmix);
    else if (a->solmix < MINVAL && b->solmix < MINVAL) mix = 0.5;
    else if (a->solmix < MINVAL) mix = 0.0;
    else mix = 1.0;
    if (a->solref[0] > 0 && b->solref[0] > 0) for (int i = 0; i < 2; i++) c->solref[i] = mix * a->solref[i] + (1 - mix) * b->solref[i];
    else for (int i = 0; i < 2; i++) c->solref[i] = a->solref[i] < b->solref[i] ? a->solref[i] : b->solref[i];
    for (int i = 0; i < 5; i++) c->solimp[i] = mix * a->solimp[i] + (1 - mix) * b->solimp[i];
    for (int i = 0; i < 3; i++) fri[i] = fa[i] > fb[i] ? fa[i] : fb[i];
  } else {
    const QsGeomParams* w = a->priority > b->priority ? a : b;
    const double* fw = a->priority > b->priority ? fa : fb;
    c->dim = w->condim;
    memcpy(c->solref, w->solref, sizeof(c->solref));
    memcpy(c->solimp, w->solimp, sizeof(c->solimp));
    memcpy(fri, fw, sizeof(fri));
  }
  for (int i = 0; i < 3; i++) if (fri[i] < MINMU) fri[i] = MINMU;
  c->friction[0] = c->friction[1] = fri[0]; c->friction[2] = fri[1]; c->friction[3] = c->friction[4] = fri[2];
  double margin = a->margin > b->margin ? a->margin : b->margin, gap = a->gap > b->gap ? a->gap : b->gap;
  c->includemargin = margin - gap;
}

static void geomFriction(const OData* d, int g, double* f) {
  const QsModel* m = &d->m;
  memcpy(f, m->geom_par[g].friction, 3 * sizeof(double));
  if (m->geom_foot_leg[g] >= 0 && d->mu_feet >= 0) { f[0] = d->mu_feet; f[1] = 0.005; f[2] = 0.0; } /* quadruped_env.py:1290-1296 */
}
static void floorFriction(const OData* d, double* f) {
  memcpy(f, d->m.floor_par.friction, 3 * sizeof(double));
  if (d->mu_floor >= 0) { f[0] = d->mu_floor; f[1] = 0.005; f[2] = 0.0; }
}

static OContact* addContact(OData* d, int g, int wgeom, double sign, double dist, const double* pos, const double* normal,
                            const double* yaxis, const QsGeomParams* wpar, const double* wfri) {
  if (d->ncon >= MAXCON) { d->overflow = 1; return NULL; }
  OContact* c = &d->con[d->ncon++];
  memset(c, 0, sizeof(*c));
  c->dist = dist; c->geom = g; c->body = d->m.geom_body[g]; c->wgeom = wgeom; c->sign = sign;
  memcpy(c->pos, pos, 3 * sizeof(double));
  memcpy(c->frame, normal, 3 * sizeof(double));
  if (yaxis) memcpy(c->frame + 3, yaxis, 3 * sizeof(double));
  makeFrame(c->frame);
  double gf[3];
  geomFriction(d, g, gf);
  mixParams(wpar, &d->m.geom_par[g], wfri, gf, c);
  c->exclude = !(dist < c->includemargin);
  return c;
}

/* plane z=0 through the origin with normal +z (scene_flat.xml:32) vs every robot geom. [MJ] mjc_PlaneSphere / PlaneCapsule /
 * PlaneBox / PlaneConvex (support vertex only: the hull graphs here are fine enough that neighbour vertices never pass
 * the engine's "far from first contact" test) */
static void collideFloor(OData* d) {
  const QsModel* m = &d->m;
  const double n[3] = {0, 0, 1};
  double wfri[3];
  floorFriction(d, wfri);
  for (int g = 0; g < m->ngeom; g++) {
    const QsGeomParams* gp = &m->geom_par[g];
    double margin = gp->margin > m->floor_par.margin ? gp->margin : m->floor_par.margin;
    const double *gx = d->geom_xpos[g], *gm = d->geom_xmat[g], *sz = m->geom_size[g];
    int b = m->geom_body[g];
    switch (m->geom_type[g]) {
      case QS_GEOM_SPHERE: {
        double dist = gx[2] - sz[0];
        if (dist > margin) break;
        double pos[3] = {gx[0], gx[1], gx[2] - (sz[0] + 0.5 * dist)};
        addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri);
      } break;
      case QS_GEOM_CAPSULE: {
        double axis[3] = {gm[2], gm[5], gm[8]};
        for (int s = 1; s >= -1; s -= 2) {
          double p[3] = {gx[0] + s * axis[0] * sz[1], gx[1] + s * axis[1] * sz[1], gx[2] + s * axis[2] * sz[1]};
          double dist = p[2] - sz[0];
          if (dist > margin) continue;
          double pos[3] = {p[0], p[1], p[2] - (sz[0] + 0.5 * dist)};
          addContact(d, g, 0, 1, dist, pos, n, axis, &m->floor_par, wfri);
        }
      } break;
      case QS_GEOM_BOX: {
        int cnt = 0;
        for (int i = 0; i < 8 && cnt < 4; i++) {
          double v[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]}, corner[3];
          mulMatVec3(corner, gm, v);
          double ldist = corner[2];
          if (gx[2] + ldist > margin || ldist > 0) continue;
          double dist = gx[2] + ldist;
          double pos[3] = {corner[0] + gx[0], corner[1] + gx[1], corner[2] + gx[2] - 0.5 * dist};
          addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri);
          cnt++;
        }
      } break;
      case QS_GEOM_CYLINDER: {
        /* [MJ] mjc_PlaneCylinder (b2.xml:96, go1.xml:27-44): the rim point nearest to the plane on the lower cap, the matching
         * point of the other cap, and two more points of the lower rim at +-120 degrees */
        double axis[3] = {gm[2], gm[5], gm[8]}, axis_z = axis[2], rim[3];
        if (axis_z > 0) { for (int i = 0; i < 3; i++) axis[i] = -axis[i]; axis_z = -axis_z; }
        const double height = gx[2];
        for (int i = 0; i < 3; i++) rim[i] = axis[i] * axis_z - n[i];
        const double rim_n2 = dot3(rim, rim);
        if (rim_n2 >= 1e-30) { const double scl = sz[0] / sqrt(rim_n2); for (int i = 0; i < 3; i++) rim[i] *= scl; }
        else { rim[0] = gm[0] * sz[0]; rim[1] = gm[3] * sz[0]; rim[2] = gm[6] * sz[0]; }
        const double rim_z = rim[2];
        double half[3] = {axis[0] * sz[1], axis[1] * sz[1], axis[2] * sz[1]};
        axis_z *= sz[1];
        double dist = height + axis_z + rim_z;
        if (dist > margin) break;
        { double pos[3]; for (int i = 0; i < 3; i++) pos[i] = gx[i] + rim[i] + half[i] - n[i] * dist * 0.5;
          addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri); }
        dist = height - axis_z + rim_z;
        if (!(dist > margin)) { double pos[3]; for (int i = 0; i < 3; i++) pos[i] = gx[i] + rim[i] - half[i] - n[i] * dist * 0.5;
          addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri); }
        dist = height + axis_z - 0.5 * rim_z;
        if (!(dist > margin)) {
          double side[3];
          cross3(side, rim, half);
          const double nv = norm3(side), scl = nv > 0 ? sz[0] * sqrt(3.0) * 0.5 / nv : 0.0;
          for (int sgn = 1; sgn >= -1; sgn -= 2) {
            double pos[3];
            for (int i = 0; i < 3; i++) pos[i] = gx[i] + sgn * scl * side[i] + half[i] - 0.5 * rim[i] - n[i] * dist * 0.5;
            addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri);
          }
        }
      } break;
      case QS_GEOM_MESH: {
        /* support vertex in -normal direction; vertices are stored in the body frame */
        double c[3], tmp[3];
        mulMatVec3(tmp, d->xmat[b], m->geom_bcenter[g]);
        for (int i = 0; i < 3; i++) c[i] = d->xpos[b][i] + tmp[i];
        if (c[2] - m->geom_rbound[g] > margin) break; /* broad phase */
        double dirl[3] = {-d->xmat[b][6], -d->xmat[b][7], -d->xmat[b][8]}; /* R^T * (-n) */
        const double* v = d->vert + 3 * m->geom_vertadr[g];
        int best = 0;
        double bv = -1e300;
        for (int i = 0; i < m->geom_vertnum[g]; i++) {
          double s = dot3(v + 3 * i, dirl);
          if (s > bv) { bv = s; best = i; }
        }
        double p[3];
        mulMatVec3(p, d->xmat[b], v + 3 * best);
        for (int i = 0; i < 3; i++) p[i] += d->xpos[b][i];
        double dist = p[2];
        if (dist > margin) break;
        double pos[3] = {p[0], p[1], p[2] - 0.5 * dist};
        addContact(d, g, 0, 1, dist, pos, n, NULL, &m->floor_par, wfri);
      } break;
      default: break;
    }
  }
}

/* ---- terrain beyond the floor plane (perlin height field, random boxes).
 * Robot geoms are reduced to "feature points" (sphere centre + radius; capsule end spheres; box corners), each tested against
 * the surface below / around it.  This is the engine's exact rule for sphere-box and plane-like cases and an approximation of
 * its capsule-box / box-box / prism-based hfield routines (no edge-edge contacts) -- [MJ-approx], documented in DESIGN.md.
 * Convex meshes use the same idea with their hull vertices as the feature points: against the height field the deepest vertex
 * (signed distance to the triangle plane under it) gives one contact per mesh; against a box the deepest vertex per (mesh, box)
 * pair, at most 4 per mesh (deepest first), the box being geom1 (box < mesh in the engine's type order). */
static int hfieldHeight(const OData* d, double x, double y, double* z, double* n) {
  const QsModel* m = &d->m;
  double sx = m->hf_size[0], sy = m->hf_size[1], sz = m->hf_size[2];
  double lx = x - m->hf_pos[0], ly = y - m->hf_pos[1];
  if (lx < -sx || lx > sx || ly < -sy || ly > sy) return 0;
  int nc = m->hf_ncol, nr = m->hf_nrow;
  double dx = 2 * sx / (nc - 1), dy = 2 * sy / (nr - 1);
  int c = (int)floor((lx + sx) / dx), r = (int)floor((ly + sy) / dy);
  if (c > nc - 2) c = nc - 2; if (r > nr - 2) r = nr - 2; if (c < 0) c = 0; if (r < 0) r = 0;
  double u = (lx + sx - c * dx) / dx, v = (ly + sy - r * dy) / dy;
  double z00 = sz * d->hf[r * nc + c], z10 = sz * d->hf[r * nc + c + 1], z01 = sz * d->hf[(r + 1) * nc + c], z11 = sz * d->hf[(r + 1) * nc + c + 1];
  double gx, gy; /* surface gradient; cells are split along the (0,0)-(1,1) diagonal [MJ] */
  if (u >= v) { *z = z00 + u * (z10 - z00) + v * (z11 - z10); gx = (z10 - z00) / dx; gy = (z11 - z10) / dy; }
  else { *z = z00 + u * (z11 - z01) + v * (z01 - z00); gx = (z11 - z01) / dx; gy = (z01 - z00) / dy; }
  *z += m->hf_pos[2];
  double inv = 1.0 / sqrt(gx * gx + gy * gy + 1);
  n[0] = -gx * inv; n[1] = -gy * inv; n[2] = inv;
  return 1;
}

static void pointBoxContact(OData* d, int g, int b, const double* p, double r, int geom_type, const double* yaxis);
static double pointBoxDistance(const QsModel* m, int b, const double* p, double* nw);

/* point feature (centre p, radius r) of robot geom g against every terrain surface except the floor plane */
static void collidePointTerrain(OData* d, int g, const double* p, double r, int geom_type, const double* yaxis) {
  const QsModel* m = &d->m;
  const QsGeomParams* gp = &m->geom_par[g];
  if (m->terrain_type == QS_TERRAIN_HFIELD) {
    double z, n[3];
    if (!hfieldHeight(d, p[0], p[1], &z, n)) return;
    double margin = gp->margin > m->hf_par.margin ? gp->margin : m->hf_par.margin;
    double dist = (p[2] - z) * n[2] - r; /* distance to the plane of the triangle under the point */
    if (dist > margin) return;
    double pos[3] = {p[0] - n[0] * (r + 0.5 * dist), p[1] - n[1] * (r + 0.5 * dist), p[2] - n[2] * (r + 0.5 * dist)};
    addContact(d, g, 1, 1, dist, pos, n, yaxis, &m->hf_par, m->hf_par.friction);
  } else if (m->terrain_type == QS_TERRAIN_BOXES) {
    for (int b = 0; b < m->nbox; b++) pointBoxContact(d, g, b, p, r, geom_type, yaxis);
  }
}

/* point feature (centre p, radius r) of robot geom g against static box b */
static void pointBoxContact(OData* d, int g, int b, const double* p, double r, int geom_type, const double* yaxis) {
  const QsModel* m = &d->m;
  const QsGeomParams* gp = &m->geom_par[g];
  {
    {
      double R[9], q[3], rel[3] = {p[0] - m->box_pos[b][0], p[1] - m->box_pos[b][1], p[2] - m->box_pos[b][2]};
      const double* h = m->box_half[b];
      if (dot3(rel, rel) > (norm3(h) + r + 0.01) * (norm3(h) + r + 0.01)) return;
      quat2Mat(R, m->box_quat[b]);
      mulMatTVec3(q, R, rel);
      double margin = gp->margin > m->box_par.margin ? gp->margin : m->box_par.margin;
      double cl[3], dl[3], nl[3], dist;
      int inside = 1;
      for (int i = 0; i < 3; i++) { cl[i] = q[i] < -h[i] ? -h[i] : (q[i] > h[i] ? h[i] : q[i]); dl[i] = q[i] - cl[i]; if (dl[i] != 0) inside = 0; }
      if (!inside) {
        double len = norm3(dl);
        dist = len - r;
        for (int i = 0; i < 3; i++) nl[i] = dl[i] / len;
      } else {
        int best = 0; double depth = 1e300;
        for (int i = 0; i < 3; i++) { double e = h[i] - fabs(q[i]); if (e < depth) { depth = e; best = i; } }
        dist = -depth - r;
        nl[0] = nl[1] = nl[2] = 0; nl[best] = q[best] >= 0 ? 1 : -1;
      }
      if (dist > margin) return;
      double nw[3];
      mulMatVec3(nw, R, nl); /* box -> point */
      double pos[3] = {p[0] - nw[0] * (r + 0.5 * dist), p[1] - nw[1] * (r + 0.5 * dist), p[2] - nw[2] * (r + 0.5 * dist)};
      /* geom1/geom2 ordered by type: sphere / capsule sort before box, so the robot geom is geom1 and the normal flips [MJ] */
      int robot_first = geom_type == QS_GEOM_SPHERE || geom_type == QS_GEOM_CAPSULE || geom_type == QS_GEOM_CYLINDER;
      double nn[3] = {robot_first ? -nw[0] : nw[0], robot_first ? -nw[1] : nw[1], robot_first ? -nw[2] : nw[2]};
      addContact(d, g, 1 + b, robot_first ? -1 : 1, dist, pos, nn, yaxis, &m->box_par, m->box_friction[b]);
    }
  }
}

/* signed distance of point p to box b (negative inside) and the outward box normal there (world frame) */
static double pointBoxDistance(const QsModel* m, int b, const double* p, double* nw) {
  double R[9], q[3], rel[3] = {p[0] - m->box_pos[b][0], p[1] - m->box_pos[b][1], p[2] - m->box_pos[b][2]};
  const double* h = m->box_half[b];
  quat2Mat(R, m->box_quat[b]);
  mulMatTVec3(q, R, rel);
  double cl[3], dl[3], nl[3] = {0, 0, 0}, dist;
  int inside = 1;
  for (int i = 0; i < 3; i++) { cl[i] = q[i] < -h[i] ? -h[i] : (q[i] > h[i] ? h[i] : q[i]); dl[i] = q[i] - cl[i]; if (dl[i] != 0) inside = 0; }
  if (!inside) {
    double len = norm3(dl);
    dist = len;
    for (int i = 0; i < 3; i++) nl[i] = dl[i] / len;
  } else {
    int best = 0; double depth = 1e300;
    for (int i = 0; i < 3; i++) { double e = h[i] - fabs(q[i]); if (e < depth) { depth = e; best = i; } }
    dist = -depth;
    nl[best] = q[best] >= 0 ? 1 : -1;
  }
  mulMatVec3(nw, R, nl);
  return dist;
}

/* Parameter t in [-L, L] of the point of the segment c + t*a (|a| = 1) nearest to static box b.  In the box frame the squared
 * distance f(t) = sum_i excess_i(t)^2 (excess = how far coordinate i lies outside its slab) is convex and C1, so its derivative
 * g(t) = sum_i dv_i * excess_i(t) is monotone: bisect for its root. */
static double segmentBoxClosest(const QsModel* m, int b, const double* c, const double* a, double L) {
  double R[9], q0[3], dv[3], rel[3] = {c[0] - m->box_pos[b][0], c[1] - m->box_pos[b][1], c[2] - m->box_pos[b][2]};
  const double* h = m->box_half[b];
  quat2Mat(R, m->box_quat[b]);
  mulMatTVec3(q0, R, rel);
  mulMatTVec3(dv, R, a);
#define SEG_G(t, out) do { double g_ = 0; for (int i_ = 0; i_ < 3; i_++) { double q_ = q0[i_] + (t) * dv[i_]; \
    if (q_ > h[i_]) g_ += dv[i_] * (q_ - h[i_]); else if (q_ < -h[i_]) g_ += dv[i_] * (q_ + h[i_]); } out = g_; } while (0)
  double lo = -L, hi = L, glo, ghi;
  SEG_G(lo, glo); SEG_G(hi, ghi);
  if (glo >= 0) return -L;
  if (ghi <= 0) return L;
  for (int it = 0; it < 60; it++) {
    double mid = 0.5 * (lo + hi), gm;
    SEG_G(mid, gm);
    if (gm < 0) lo = mid; else hi = mid;
  }
#undef SEG_G
  return 0.5 * (lo + hi);
}

/* Capsule against static boxes beyond its two end spheres: where an interior point of the axis is strictly nearer to a box than
 * both ends (a leg lying across a stair edge), that point is a third sphere feature.  Not generated when the ends are as near
 * (capsule flat on a face: the two end contacts carry it). [MJ-approx of mjc_CapsuleBox] */
static void collideCapsuleMidBoxes(OData* d, int g, const double* c, const double* a, double L, double r) {
  const QsModel* m = &d->m;
  for (int b = 0; b < m->nbox; b++) {
    double rel[3] = {c[0] - m->box_pos[b][0], c[1] - m->box_pos[b][1], c[2] - m->box_pos[b][2]};
    double reach = norm3(m->box_half[b]) + L + r + 0.01;
    if (dot3(rel, rel) > reach * reach) continue;
    double t = segmentBoxClosest(m, b, c, a, L);
    if (!(t > -L * (1 - 1e-6) && t < L * (1 - 1e-6))) continue;
    double pm[3] = {c[0] + t * a[0], c[1] + t * a[1], c[2] + t * a[2]}, p1[3], p2[3], nw[3];
    for (int i = 0; i < 3; i++) { p1[i] = c[i] + L * a[i]; p2[i] = c[i] - L * a[i]; }
    double dm = pointBoxDistance(m, b, pm, nw), d1 = pointBoxDistance(m, b, p1, nw), d2 = pointBoxDistance(m, b, p2, nw);
    if (!(dm < fmin(d1, d2) - 1e-6)) continue;
    pointBoxContact(d, g, b, pm, r, QS_GEOM_CAPSULE, a);
  }
}

/* convex mesh g (hull vertices in the body frame) against the height field / the static boxes */
static void collideMeshTerrain(OData* d, int g) {
  const QsModel* m = &d->m;
  const QsGeomParams* gp = &m->geom_par[g];
  const int b = m->geom_body[g], nv = m->geom_vertnum[g];
  const double* v = d->vert + 3 * m->geom_vertadr[g];
  if (m->terrain_type == QS_TERRAIN_HFIELD) {
    double margin = gp->margin > m->hf_par.margin ? gp->margin : m->hf_par.margin;
    double best = 1e300, bp[3] = {0, 0, 0}, bn[3] = {0, 0, 1};
    for (int i = 0; i < nv; i++) {
      double p[3], z, n[3];
      mulMatVec3(p, d->xmat[b], v + 3 * i);
      for (int k = 0; k < 3; k++) p[k] += d->xpos[b][k];
      if (!hfieldHeight(d, p[0], p[1], &z, n)) continue;
      double dist = (p[2] - z) * n[2];
      if (dist < best) { best = dist; memcpy(bp, p, sizeof(bp)); memcpy(bn, n, sizeof(bn)); }
    }
    if (best > margin) return;
    double pos[3] = {bp[0] - bn[0] * 0.5 * best, bp[1] - bn[1] * 0.5 * best, bp[2] - bn[2] * 0.5 * best};
    addContact(d, g, 1, 1, best, pos, bn, NULL, &m->hf_par, m->hf_par.friction);
  } else if (m->terrain_type == QS_TERRAIN_BOXES) {
    double margin = gp->margin > m->box_par.margin ? gp->margin : m->box_par.margin;
    double c[3], tmp[3];
    mulMatVec3(tmp, d->xmat[b], m->geom_bcenter[g]);
    for (int k = 0; k < 3; k++) c[k] = d->xpos[b][k] + tmp[k];
    for (int bx = 0; bx < m->nbox; bx++) {
      double rel[3] = {c[0] - m->box_pos[bx][0], c[1] - m->box_pos[bx][1], c[2] - m->box_pos[bx][2]};
      double reach = norm3(m->box_half[bx]) + m->geom_rbound[g] + margin + 0.01;
      if (dot3(rel, rel) > reach * reach) continue; /* bounding spheres: conservative */
      double best = 1e300, bp[3] = {0, 0, 0}, bn[3] = {0, 0, 1};
      for (int i = 0; i < nv; i++) {
        double p[3], nw[3];
        mulMatVec3(p, d->xmat[b], v + 3 * i);
        for (int k = 0; k < 3; k++) p[k] += d->xpos[b][k];
        double dist = pointBoxDistance(m, bx, p, nw);
        if (dist < best) { best = dist; memcpy(bp, p, sizeof(bp)); memcpy(bn, nw, sizeof(bn)); }
      }
      if (best > margin) continue;
      double pos[3] = {bp[0] - bn[0] * 0.5 * best, bp[1] - bn[1] * 0.5 * best, bp[2] - bn[2] * 0.5 * best};
      addContact(d, g, 1 + bx, 1, best, pos, bn, NULL, &m->box_par, m->box_friction[bx]);
    }
  }
}

static void collideTerrain(OData* d) {
  const QsModel* m = &d->m;
  if (m->terrain_type == QS_TERRAIN_FLAT) return;
  for (int g = 0; g < m->ngeom; g++) {
    const double *gx = d->geom_xpos[g], *gm = d->geom_xmat[g], *sz = m->geom_size[g];
    int before = d->ncon;
    switch (m->geom_type[g]) {
      case QS_GEOM_SPHERE: collidePointTerrain(d, g, gx, sz[0], QS_GEOM_SPHERE, NULL); break;
      case QS_GEOM_CAPSULE: {
        double axis[3] = {gm[2], gm[5], gm[8]};
        for (int s = 1; s >= -1; s -= 2) {
          double p[3] = {gx[0] + s * axis[0] * sz[1], gx[1] + s * axis[1] * sz[1], gx[2] + s * axis[2] * sz[1]};
          collidePointTerrain(d, g, p, sz[0], QS_GEOM_CAPSULE, axis);
        }
        if (m->terrain_type == QS_TERRAIN_BOXES) collideCapsuleMidBoxes(d, g, gx, axis, sz[1], sz[0]);
      } break;
      case QS_GEOM_BOX:
        for (int i = 0; i < 8; i++) {
          double v[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]}, c[3];
          mulMatVec3(c, gm, v);
          double p[3] = {c[0] + gx[0], c[1] + gx[1], c[2] + gx[2]};
          collidePointTerrain(d, g, p, 0.0, QS_GEOM_BOX, NULL);
        }
        break;
      case QS_GEOM_CYLINDER:  /* eight rim points (four per cap) as point features: an approximation, like the box corners */
        for (int i = 0; i < 8; i++) {
          const double a = (i & 1) ? sz[0] : -sz[0];
          double v[3] = {(i & 2) ? a : 0.0, (i & 2) ? 0.0 : a, (i & 4) ? sz[1] : -sz[1]}, c[3];
          mulMatVec3(c, gm, v);
          double p[3] = {c[0] + gx[0], c[1] + gx[1], c[2] + gx[2]};
          collidePointTerrain(d, g, p, 0.0, QS_GEOM_CYLINDER, NULL);
        }
        break;
      case QS_GEOM_MESH: collideMeshTerrain(d, g); break;
      default: break;
    }
    /* at most 4 terrain contacts per geom, deepest first (the engine caps its multi-contact routines similarly) */
    int cnt = d->ncon - before;
    if (cnt > 4) {
      OContact* c = d->con + before;
      for (int i = 0; i < cnt; i++) for (int j = i + 1; j < cnt; j++) if (c[j].dist < c[i].dist) { OContact t = c[i]; c[i] = c[j]; c[j] = t; }
      d->ncon = before + 4;
    }
  }
}

static void collision(OData* d) {
  d->ncon = 0;
  d->overflow = 0;
  collideFloor(d);
  collideTerrain(d);
}

/* downward ray from `org` against the static terrain: distance to the nearest hit, -1 if none. [MJ] mj_ray with
 * geomgroup {0,4,5}, flg_static=1 (sensors/heightmap.py:77-99) */
static double rayDown(const OData* d, const double* org) {
  const QsModel* m = &d->m;
  double best = -1;
  if (org[2] >= 0) best = org[2]; /* floor plane z = 0 */
  if (m->terrain_type == QS_TERRAIN_HFIELD) {
    double z, n[3];
    if (hfieldHeight(d, org[0], org[1], &z, n) && org[2] >= z) { double t = org[2] - z; if (best < 0 || t < best) best = t; }
  } else if (m->terrain_type == QS_TERRAIN_BOXES) {
    for (int b = 0; b < m->nbox; b++) {
      double R[9], o[3], dl[3], rel[3] = {org[0] - m->box_pos[b][0], org[1] - m->box_pos[b][1], org[2] - m->box_pos[b][2]}, dw[3] = {0, 0, -1};
      quat2Mat(R, m->box_quat[b]);
      mulMatTVec3(o, R, rel);
      mulMatTVec3(dl, R, dw);
      const double* h = m->box_half[b];
      double tmin = -1e300, tmax = 1e300;
      int miss = 0;
      for (int i = 0; i < 3; i++) {
        if (fabs(dl[i]) < 1e-12) { if (o[i] < -h[i] || o[i] > h[i]) miss = 1; continue; }
        double t1 = (-h[i] - o[i]) / dl[i], t2 = (h[i] - o[i]) / dl[i];
        if (t1 > t2) { double t = t1; t1 = t2; t2 = t; }
        if (t1 > tmin) tmin = t1;
        if (t2 < tmax) tmax = t2;
      }
      if (miss || tmin > tmax || tmax < 0) continue;
      double t = tmin >= 0 ? tmin : tmax; /* origin inside the box: the exit face */
      if (best < 0 || t < best) best = t;
    }
  }
  return best;
}

/* ------------------------------------------------------------------ constraints */
/* [MJ] getimpedance, SURVEY App. A.6 */
static double impedance(const double* solimp_in, double x /* pos - margin */) {
  double s[5];
  memcpy(s, solimp_in, sizeof(s));
  for (int i = 0; i < 2; i++) { if (s[i] < MINIMP) s[i] = MINIMP; if (s[i] > MAXIMP) s[i] = MAXIMP; }
  if (s[2] < 0) s[2] = 0;
  if (s[3] < MINIMP) s[3] = MINIMP; if (s[3] > MAXIMP) s[3] = MAXIMP;
  if (s[4] < 1) s[4] = 1;
  if (s[0] == s[1] || s[2] <= MINVAL) return 0.5 * (s[0] + s[1]);
  x = fabs(x) / s[2];
  if (x >= 1) return s[1];
  if (x == 0) return s[0];
  double y;
  if (s[4] == 1) y = x;
  else if (x <= s[3]) y = pow(x, s[4]) / pow(s[3], s[4] - 1);
  else y = 1 - pow(1 - x, s[4]) / pow(1 - s[3], s[4] - 1);
  return s[0] + y * (s[1] - s[0]);
}

static int addRow(OData* d, int type, int id, double pos, double margin, double floss) {
  int r = d->nefc++;
  memset(d->efc_J[r], 0, sizeof(d->efc_J[r]));
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_floss[r] = floss;
  return r;
}

/* [MJ] mj_makeConstraint + mj_makeImpedance + mj_referenceConstraint, SURVEY App. A.6 */
static void makeConstraint(OData* d) {
  const QsModel* m = &d->m;
  d->nefc = 0;
  /* 1. dof friction loss */
  for (int i = 0; i < NV; i++)
    if (m->dof_frictionloss[i] > 0) {
      int r = addRow(d, T_FRICTION, i, 0, 0, m->dof_frictionloss[i]);
      d->efc_J[r][i] = 1;
    }
  /* 2. joint limits */
  for (int j = 0; j < QS_NJNT; j++)
    if (m->jnt_limited[j]) {
      double value = d->qpos[7 + j];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
        if (dist < m->jnt_margin[j]) {
          int r = addRow(d, T_LIMIT, j, dist, m->jnt_margin[j], 0);
          d->efc_J[r][6 + j] = -side;
        }
      }
    }
  /* 3. contacts */
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact* c = &d->con[ci];
    c->efc_address = -1;
    if (c->exclude) continue;
    double jp[3][NV], jr[3][NV], Jc[6][NV];
    jacPoint(d, jp, jr, c->pos, c->body);
    for (int k = 0; k < 3; k++)
      for (int v = 0; v < NV; v++) {
        Jc[k][v] = c->sign * (c->frame[3 * k] * jp[0][v] + c->frame[3 * k + 1] * jp[1][v] + c->frame[3 * k + 2] * jp[2][v]);
        Jc[3 + k][v] = c->sign * (c->frame[3 * k] * jr[0][v] + c->frame[3 * k + 1] * jr[1][v] + c->frame[3 * k + 2] * jr[2][v]);
      }
    c->efc_address = d->nefc;
    if (c->dim == 1) {
      int r = addRow(d, T_CONTACT_FRICTIONLESS, ci, c->dist, c->includemargin, 0);
      memcpy(d->efc_J[r], Jc[0], sizeof(Jc[0]));
    } else if (m->cone == QS_CONE_PYRAMIDAL) {
      for (int k = 1; k < c->dim; k++)
        for (int s = 1; s >= -1; s -= 2) {
          int r = addRow(d, T_CONTACT_PYRAMIDAL, ci, c->dist, c->includemargin, 0);
          for (int v = 0; v < NV; v++) d->efc_J[r][v] = Jc[0][v] + s * c->friction[k - 1] * Jc[k][v];
        }
    } else {
      for (int k = 0; k < c->dim; k++) {
        int r = addRow(d, T_CONTACT_ELLIPTIC, ci, k == 0 ? c->dist : 0, k == 0 ? c->includemargin : 0, 0);
        memcpy(d->efc_J[r], Jc[k], sizeof(Jc[k]));
      }
    }
  }
  /* diagApprox, impedance, R, D, aref */
  double h = m->timestep;
  for (int r = 0; r < d->nefc;) {
    int type = d->efc_type[r], id = d->efc_id[r], nrow = 1;
    const double *solref, *solimp;
    if (type == T_FRICTION) { d->efc_diagApprox[r] = m->dof_invweight0[id]; solref = m->dof_solref[id]; solimp = m->dof_solimp[id]; }
    else if (type == T_LIMIT) { d->efc_diagApprox[r] = m->dof_invweight0[6 + id]; solref = m->jnt_solref[id]; solimp = m->jnt_solimp[id]; }
    else {
      OContact* c = &d->con[id];
      double tran = m->body_invweight0[c->body][0], rot = m->body_invweight0[c->body][1]; /* world body adds 0 */
      solref = c->solref; solimp = c->solimp;
      if (type == T_CONTACT_FRICTIONLESS) d->efc_diagApprox[r] = tran;
      else if (type == T_CONTACT_ELLIPTIC) { nrow = c->dim; for (int k = 0; k < nrow; k++) d->efc_diagApprox[r + k] = k < 3 ? tran : rot; }
      else { nrow = 2 * (c->dim - 1); for (int k = 0; k < nrow; k++) { double f = c->friction[k / 2]; d->efc_diagApprox[r + k] = tran + f * f * (k / 2 < 2 ? tran : rot); } }
    }
    /* [MJ] getsolparam: standard (positive) solref only */
    double dmax = solimp[1] < MINIMP ? MINIMP : (solimp[1] > MAXIMP ? MAXIMP : solimp[1]);
    double tc = solref[0] > 2 * h ? solref[0] : 2 * h, dr = solref[1];
    double K = 1 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr), B = 2 / fmax(MINVAL, dmax * tc);
    for (int k = 0; k < nrow; k++) {
      int q = r + k;
      double imp = impedance(solimp, d->efc_pos[q] - d->efc_margin[q]);
      d->efc_imp[q] = imp;
      d->efc_R[q] = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[q] / imp);
      double Kq = K;
      if (type == T_FRICTION || (type == T_CONTACT_ELLIPTIC && k > 0)) Kq = 0;
      double vel = 0;
      for (int v = 0; v < NV; v++) vel += d->efc_J[q][v] * d->qvel[v];
      d->efc_vel[q] = vel;
      d->efc_aref[q] = -B * vel - Kq * imp * (d->efc_pos[q] - d->efc_margin[q]);
    }
    /* friction-cone adjustment of R */
    if (type == T_CONTACT_ELLIPTIC && nrow > 1) {
      OContact* c = &d->con[id];
      d->efc_R[r + 1] = d->efc_R[r] / fmax(MINVAL, m->impratio);
      c->mu = c->friction[0] * sqrt(d->efc_R[r + 1] / d->efc_R[r]);
      for (int k = 2; k < nrow; k++) d->efc_R[r + k] = d->efc_R[r + 1] * c->friction[0] * c->friction[0] / (c->friction[k - 1] * c->friction[k - 1]);
    } else if (type == T_CONTACT_PYRAMIDAL) {
      OContact* c = &d->con[id];
      c->mu = c->friction[0] * sqrt(1 / fmax(MINVAL, m->impratio));
      double Rpy = 2 * c->mu * c->mu * d->efc_R[r];
      for (int k = 0; k < nrow; k++) d->efc_R[r + k] = Rpy;
    }
    for (int k = 0; k < nrow; k++) d->efc_D[r + k] = 1 / d->efc_R[r + k];
    r += nrow;
  }
}

/* ------------------------------------------------------------------ velocity / force stage */
/* [MJ] mj_comVel */
static void comVel(OData* d) {
  memset(d->cvel[0], 0, sizeof(d->cvel[0]));
  /* base: translations first (cdof_dot = 0), then rotations using the velocity after translations */
  double v[6] = {0, 0, 0, 0, 0, 0};
  memset(d->cdof_dot, 0, sizeof(d->cdof_dot));
  for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) v[i] += d->cdof[k][i] * d->qvel[k];
  for (int k = 3; k < 6; k++) crossMotion(d->cdof_dot[k], v, d->cdof[k]);
  for (int k = 3; k < 6; k++) for (int i = 0; i < 6; i++) v[i] += d->cdof[k][i] * d->qvel[k];
  memcpy(d->cvel[1], v, sizeof(v));
  for (int b = 2; b < NB; b++) {
    int p = d->m.body_parent[b], k = 6 + b - 2;
    memcpy(v, d->cvel[p], sizeof(v));
    crossMotion(d->cdof_dot[k], v, d->cdof[k]);
    for (int i = 0; i < 6; i++) v[i] += d->cdof[k][i] * d->qvel[k];
    memcpy(d->cvel[b], v, sizeof(v));
  }
}

/* [MJ] mj_rne; with_acc adds cdof*qacc (mj_rnePostConstraint uses it for cacc) */
static void rne(OData* d, int with_acc, double* result) {
  const QsModel* m = &d->m;
  memset(d->cacc[0], 0, sizeof(d->cacc[0]));
  for (int i = 0; i < 3; i++) d->cacc[0][3 + i] = -m->gravity[i];
  memset(d->cfrc[0], 0, sizeof(d->cfrc[0]));
  for (int b = 1; b < NB; b++) {
    int p = m->body_parent[b], d0 = (b == 1) ? 0 : 6 + b - 2, nd = (b == 1) ? 6 : 1;
    double a[6], t1[6], t2[6];
    memcpy(a, d->cacc[p], sizeof(a));
    for (int k = d0; k < d0 + nd; k++)
      for (int i = 0; i < 6; i++) a[i] += d->cdof_dot[k][i] * d->qvel[k] + (with_acc ? d->cdof[k][i] * d->qacc[k] : 0);
    memcpy(d->cacc[b], a, sizeof(a));
    mulInertVec(t1, d->cinert[b], a);
    mulInertVec(t2, d->cinert[b], d->cvel[b]);
    crossForce(d->cfrc[b], d->cvel[b], t2);
    for (int i = 0; i < 6; i++) d->cfrc[b][i] += t1[i];
  }
  if (!result) return;
  for (int b = NB - 1; b > 1; b--) {
    int p = m->body_parent[b];
    for (int i = 0; i < 6; i++) d->cfrc[p][i] += d->cfrc[b][i];
  }
  for (int k = 0; k < NV; k++) {
    double s = 0;
    for (int i = 0; i < 6; i++) s += d->cdof[k][i] * d->cfrc[dof_body(k)][i];
    result[k] = s;
  }
}

static void fwdSmooth(OData* d) {
  const QsModel* m = &d->m;
  comVel(d);
  rne(d, 0, d->qfrc_bias);
  for (int k = 0; k < NV; k++) d->qfrc_passive[k] = -m->dof_damping[k] * d->qvel[k];
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int a = 0; a < NU; a++) { /* [MJ] mj_fwdActuation: motor, gear 1; ctrl clamped to ctrlrange, force to forcerange */
    double c = d->ctrl[a];
    if (m->act_ctrllimited[a]) c = fmin(fmax(c, m->act_ctrlrange[a][0]), m->act_ctrlrange[a][1]);
    if (m->act_forcelimited[a]) c = fmin(fmax(c, m->act_forcerange[a][0]), m->act_forcerange[a][1]);
    d->qfrc_actuator[6 + a] = c;
  }
  for (int k = 0; k < NV; k++) d->qfrc_smooth[k] = d->qfrc_passive[k] - d->qfrc_bias[k] + d->qfrc_applied[k] + d->qfrc_actuator[k];
  cholSolve(d->L, d->qacc_smooth, d->qfrc_smooth, NV);
}

/* ------------------------------------------------------------------ solver */
typedef struct { double cost, d1, d2; } LsPoint;

/* cost of the constraint rows at jar (+ forces, states); cone Hessian blocks (dim x dim per contact) on request.
 * [MJ] mj_constraintUpdate, SURVEY App. A.7 */
static double constraintUpdate(OData* d, const double* jar, double* force, int* state, double (*coneH)[36]) {
  double cost = 0;
  for (int r = 0; r < d->nefc;) {
    int type = d->efc_type[r];
    double D = d->efc_D[r], R = d->efc_R[r];
    if (type == T_FRICTION) {
      double f = d->efc_floss[r], rf = R * f;
      if (jar[r] <= -rf) { force[r] = f; state[r] = S_LINEARNEG; cost += -f * (0.5 * rf + jar[r]); }
      else if (jar[r] >= rf) { force[r] = -f; state[r] = S_LINEARPOS; cost += -f * (0.5 * rf - jar[r]); }
      else { force[r] = -D * jar[r]; state[r] = S_QUADRATIC; cost += 0.5 * D * jar[r] * jar[r]; }
      r++;
    } else if (type != T_CONTACT_ELLIPTIC) {
      if (jar[r] >= 0) { force[r] = 0; state[r] = S_SATISFIED; }
      else { force[r] = -D * jar[r]; state[r] = S_QUADRATIC; cost += 0.5 * D * jar[r] * jar[r]; }
      r++;
    } else {
      OContact* c = &d->con[d->efc_id[r]];
      int dim = c->dim;
      double mu = c->mu, U[6], N, T = 0;
      U[0] = jar[r] * mu;
      for (int k = 1; k < dim; k++) { U[k] = jar[r + k] * c->friction[k - 1]; T += U[k] * U[k]; }
      N = U[0]; T = sqrt(T);
      if (coneH) memset(coneH[d->efc_id[r]], 0, 36 * sizeof(double));
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        for (int k = 0; k < dim; k++) { force[r + k] = 0; state[r + k] = S_SATISFIED; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int k = 0; k < dim; k++) { force[r + k] = -d->efc_D[r + k] * jar[r + k]; state[r + k] = S_QUADRATIC; cost += 0.5 * d->efc_D[r + k] * jar[r + k] * jar[r + k]; }
      } else {
        double Dm = D / fmax(MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        force[r] = -Dm * NmT * mu;
        for (int k = 1; k < dim; k++) force[r + k] = -force[r] / T * U[k] * c->friction[k - 1];
        for (int k = 0; k < dim; k++) state[r + k] = S_CONE;
        if (coneH) { /* d2/djar2 of 0.5*Dm*(N - mu*T)^2 */
          double* H = coneH[d->efc_id[r]];
          double de[6];
          de[0] = mu;
          for (int k = 1; k < dim; k++) de[k] = -mu * c->friction[k - 1] * U[k] / T;
          for (int a = 0; a < dim; a++)
            for (int b = 0; b < dim; b++) {
              double h2 = Dm * de[a] * de[b];
              if (a > 0 && b > 0) {
                double fa = c->friction[a - 1], fb = c->friction[b - 1];
                h2 += Dm * NmT * (-mu) * fa * fb * ((a == b ? 1.0 / T : 0.0) - U[a] * U[b] / (T * T * T));
              }
              H[6 * a + b] = h2;
            }
        }
      }
      r += dim;
    }
  }
  return cost;
}

typedef struct {
  OData* d;
  double Ma[NV], jar[MAXEFC], Mv[NV], Jv[MAXEFC], search[NV], grad[NV], Mgrad[NV];
  double gauss, cost;
  double quadGauss[3];
  double coneH[MAXCON][36];
  int nls;
} Ctx;

static void ctxEval(Ctx* c, int flg_cone) {
  OData* d = c->d;
  for (int i = 0; i < NV; i++) { double s = 0; for (int j = 0; j < NV; j++) s += d->M[i][j] * d->qacc[j]; c->Ma[i] = s; }
  for (int r = 0; r < d->nefc; r++) { double s = 0; for (int j = 0; j < NV; j++) s += d->efc_J[r][j] * d->qacc[j]; c->jar[r] = s - d->efc_aref[r]; }
  double cc = constraintUpdate(d, c->jar, d->efc_force, d->efc_state, flg_cone ? c->coneH : NULL);
  double g = 0;
  for (int i = 0; i < NV; i++) g += 0.5 * (c->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  c->gauss = g;
  c->cost = g + cc;
  for (int i = 0; i < NV; i++) {
    double s = 0;
    for (int r = 0; r < d->nefc; r++) s += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = s;
    c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - s;
  }
}

/* Newton direction: H = M + J^T diag(D_active) J + cone blocks; Mgrad = H^-1 grad */
static void newtonDirection(Ctx* c) {
  OData* d = c->d;
  double H[NV][NV], LH[NV][NV];
  memcpy(H, d->M, sizeof(H));
  for (int r = 0; r < d->nefc;) {
    int st = d->efc_state[r];
    if (st == S_QUADRATIC) {
      for (int i = 0; i < NV; i++) { double a = d->efc_D[r] * d->efc_J[r][i]; if (a != 0) for (int j = 0; j < NV; j++) H[i][j] += a * d->efc_J[r][j]; }
      r++;
    } else if (st == S_CONE) {
      OContact* cn = &d->con[d->efc_id[r]];
      int dim = cn->dim;
      const double* Hc = c->coneH[d->efc_id[r]];
      for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
          double h = Hc[6 * a + b];
          if (h == 0) continue;
          for (int i = 0; i < NV; i++) { double t = h * d->efc_J[r + a][i]; if (t != 0) for (int j = 0; j < NV; j++) H[i][j] += t * d->efc_J[r + b][j]; }
        }
      r += dim;
    } else r++;
  }
  if (cholFactor(LH, H, NV)) { memcpy(c->Mgrad, c->grad, sizeof(c->grad)); return; }
  cholSolve(LH, c->Mgrad, c->grad, NV);
}

/* cost and its first two derivatives along qacc + alpha*search */
static LsPoint lsEval(Ctx* c, double alpha) {
  OData* d = c->d;
  LsPoint p;
  p.cost = alpha * alpha * c->quadGauss[2] + alpha * c->quadGauss[1] + c->quadGauss[0];
  p.d1 = 2 * alpha * c->quadGauss[2] + c->quadGauss[1];
  p.d2 = 2 * c->quadGauss[2];
  c->nls++;
  for (int r = 0; r < d->nefc;) {
    int type = d->efc_type[r];
    double x = c->jar[r] + alpha * c->Jv[r], D = d->efc_D[r], jv = c->Jv[r];
    if (type == T_FRICTION) {
      double f = d->efc_floss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) { p.cost += -f * (0.5 * rf + x); p.d1 += -f * jv; }
      else if (x >= rf) { p.cost += -f * (0.5 * rf - x); p.d1 += f * jv; }
      else { p.cost += 0.5 * D * x * x; p.d1 += D * x * jv; p.d2 += D * jv * jv; }
      r++;
    } else if (type != T_CONTACT_ELLIPTIC) {
      if (x < 0) { p.cost += 0.5 * D * x * x; p.d1 += D * x * jv; p.d2 += D * jv * jv; }
      r++;
    } else {
      OContact* cn = &d->con[d->efc_id[r]];
      int dim = cn->dim;
      double mu = cn->mu, U[6], V[6], T2 = 0, UV = 0, VV = 0;
      U[0] = x * mu; V[0] = jv * mu;
      for (int k = 1; k < dim; k++) {
        U[k] = (c->jar[r + k] + alpha * c->Jv[r + k]) * cn->friction[k - 1];
        V[k] = c->Jv[r + k] * cn->friction[k - 1];
        T2 += U[k] * U[k]; UV += U[k] * V[k]; VV += V[k] * V[k];
      }
      double N = U[0], T = sqrt(T2);
      if (N >= mu * T || (T <= 0 && N >= 0)) { /* nothing */ }
      else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int k = 0; k < dim; k++) { double xk = c->jar[r + k] + alpha * c->Jv[r + k], Dk = d->efc_D[r + k], jk = c->Jv[r + k]; p.cost += 0.5 * Dk * xk * xk; p.d1 += Dk * xk * jk; p.d2 += Dk * jk * jk; }
      } else {
        double Dm = D / fmax(MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        double N1 = V[0], T1 = UV / T, T2d = VV / T - UV * UV / (T * T * T);
        p.cost += 0.5 * Dm * NmT * NmT;
        p.d1 += Dm * NmT * (N1 - mu * T1);
        p.d2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2d));
      }
      r += dim;
    }
  }
  return p;
}

/* [MJ] PrimalSearch: exact line search by safeguarded 1-D Newton with bracketing */
static double lineSearch(Ctx* c, double tolerance, double ls_tolerance, int ls_iterations, double scale) {
  OData* d = c->d;
  double snorm = 0;
  for (int i = 0; i < NV; i++) snorm += c->search[i] * c->search[i];
  snorm = sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double gtol = tolerance * ls_tolerance * snorm / scale;
  /* prepare */
  for (int i = 0; i < NV; i++) { double s = 0; for (int j = 0; j < NV; j++) s += d->M[i][j] * c->search[j]; c->Mv[i] = s; }
  for (int r = 0; r < d->nefc; r++) { double s = 0; for (int j = 0; j < NV; j++) s += d->efc_J[r][j] * c->search[j]; c->Jv[r] = s; }
  c->quadGauss[0] = c->gauss; c->quadGauss[1] = 0; c->quadGauss[2] = 0;
  for (int i = 0; i < NV; i++) { c->quadGauss[1] += c->search[i] * (c->Ma[i] - d->qfrc_smooth[i]); c->quadGauss[2] += 0.5 * c->search[i] * c->Mv[i]; }
  LsPoint p0 = lsEval(c, 0);
  double a0 = 0, a1 = a0 - p0.d1 / p0.d2;
  LsPoint p1 = lsEval(c, a1);
  if (p0.cost < p1.cost) { p1 = p0; a1 = a0; }
  if (fabs(p1.d1) < gtol) return a1;
  int dir = p1.d1 < 0 ? 1 : -1, it = 0;
  LsPoint p2 = p1; double a2 = a1;
  while (p1.d1 * dir <= -gtol && it < ls_iterations) {
    p2 = p1; a2 = a1;
    a1 = a1 - p1.d1 / p1.d2;
    p1 = lsEval(c, a1);
    it++;
    if (fabs(p1.d1) < gtol) return a1;
  }
  if (it >= ls_iterations) return a1;
  /* bracketed between (a2: derivative on the starting side) and (a1: overshoot); refine with Newton from both ends + midpoint */
  double lo = a2, hi = a1; LsPoint plo = p2, phi = p1;
  while (it < ls_iterations) {
    double cand[3] = {lo - plo.d1 / plo.d2, hi - phi.d1 / phi.d2, 0.5 * (lo + hi)};
    int moved = 0;
    for (int k = 0; k < 3; k++) {
      double a = cand[k];
      if (!((a > lo && a < hi) || (a < lo && a > hi))) continue;
      LsPoint p = lsEval(c, a);
      it++;
      if (fabs(p.d1) < gtol) return a;
      if ((p.d1 < 0) == (plo.d1 < 0)) { lo = a; plo = p; } else { hi = a; phi = p; }
      moved = 1;
    }
    if (!moved) break;
  }
  return plo.cost < phi.cost ? lo : hi;
}

/* [MJ] mj_fwdConstraint + mj_solNewton (mj_solPrimal, flg_Newton) */
static void solve(OData* d) {
  const QsModel* m = &d->m;
  d->solver_iter = 0;
  if (d->nefc == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc)); memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint)); return; }
  Ctx* c = (Ctx*)d->ctx;
  c->d = d; c->nls = 0;
  /* warm start: pick the cheaper of qacc_warmstart and qacc_smooth */
  memcpy(d->qacc, d->qacc_warmstart, sizeof(d->qacc));
  ctxEval(c, 0);
  double cost_warm = c->cost;
  memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc));
  ctxEval(c, 0);
  if (cost_warm < c->cost) memcpy(d->qacc, d->qacc_warmstart, sizeof(d->qacc));
  ctxEval(c, 1);
  double scale = 1 / (m->meaninertia * NV);
  newtonDirection(c);
  for (int i = 0; i < NV; i++) c->search[i] = -c->Mgrad[i];
  int iter = 0;
  while (iter < m->iterations) {
    double alpha = lineSearch(c, m->tolerance, m->ls_tolerance, m->ls_iterations, scale);
    if (alpha == 0) break;
    for (int i = 0; i < NV; i++) d->qacc[i] += alpha * c->search[i];
    double oldcost = c->cost;
    ctxEval(c, 1);
    newtonDirection(c);
    iter++;
    double gn = 0;
    for (int i = 0; i < NV; i++) gn += c->grad[i] * c->grad[i];
    double improvement = scale * (oldcost - c->cost), gradient = scale * sqrt(gn);
    if (improvement < m->tolerance || gradient < m->tolerance) break;
    for (int i = 0; i < NV; i++) c->search[i] = -c->Mgrad[i];
  }
  d->solver_iter = iter;
}

/* contact-frame forces, [MJ] mj_contactForce / mju_decodePyramid */
static void contactForces(OData* d) {
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact* c = &d->con[ci];
    memset(c->force, 0, sizeof(c->force));
    if (c->efc_address < 0) continue;
    const double* f = d->efc_force + c->efc_address;
    if (c->dim == 1) c->force[0] = f[0];
    else if (d->m.cone == QS_CONE_ELLIPTIC) for (int k = 0; k < c->dim; k++) c->force[k] = f[k];
    else {
      for (int k = 0; k < 2 * (c->dim - 1); k++) c->force[0] += f[k];
      for (int k = 1; k < c->dim; k++) c->force[k] = (f[2 * (k - 1)] - f[2 * (k - 1) + 1]) * c->friction[k - 1];
    }
  }
}

/* accelerometer + gyro at the IMU site, [MJ] mj_rnePostConstraint + mj_objectAcceleration, SURVEY App. A.8 */
static void sensors(OData* d) {
  const QsModel* m = &d->m;
  if (!m->has_imu) { memset(d->sensor_acc, 0, sizeof(d->sensor_acc)); memset(d->sensor_gyro, 0, sizeof(d->sensor_gyro)); return; }
  rne(d, 1, NULL); /* cacc with qacc */
  double spos[3], tmp[3], q[4], R[9];
  mulMatVec3(tmp, d->xmat[1], m->imu_pos);
  for (int i = 0; i < 3; i++) spos[i] = d->xpos[1][i] + tmp[i];
  quatMul(q, d->xquat[1], m->imu_quat);
  quat2Mat(R, q);
  double dif[3], vl[3], al[3], c1[3];
  for (int i = 0; i < 3; i++) dif[i] = spos[i] - d->com[i];
  cross3(c1, d->cvel[1], dif);
  for (int i = 0; i < 3; i++) vl[i] = d->cvel[1][3 + i] + c1[i]; /* site linear velocity */
  cross3(c1, d->cacc[1], dif);
  for (int i = 0; i < 3; i++) al[i] = d->cacc[1][3 + i] + c1[i];
  cross3(c1, d->cvel[1], vl); /* omega x v correction */
  for (int i = 0; i < 3; i++) al[i] += c1[i];
  mulMatTVec3(d->sensor_acc, R, al);
  mulMatTVec3(d->sensor_gyro, R, d->cvel[1]);
}

static void forward(OData* d) {
  kinematics(d);
  comPos(d);
  crb(d);
  cholFactor(d->L, d->M, NV);
  collision(d);
  makeConstraint(d);
  fwdSmooth(d);
  solve(d);
  contactForces(d);
  sensors(d);
}

/* [MJ] mj_Euler with implicit joint damping + mj_integratePos, SURVEY App. A.9 */
static void euler(OData* d) {
  const QsModel* m = &d->m;
  double h = m->timestep, qacc[NV];
  int damped = 0;
  for (int i = 0; i < NV; i++) if (m->dof_damping[i] > 0) damped = 1;
  if (damped) {
    double A[NV][NV], LA[NV][NV];
    double rhs[NV];
    memcpy(A, d->M, sizeof(A));
    for (int i = 0; i < NV; i++) { A[i][i] += h * m->dof_damping[i]; rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    cholFactor(LA, A, NV);
    cholSolve(LA, qacc, rhs, NV);
  } else memcpy(qacc, d->qacc, sizeof(qacc));
  for (int i = 0; i < NV; i++) d->qvel[i] += h * qacc[i];
  for (int i = 0; i < 3; i++) d->qpos[i] += h * d->qvel[i];
  double w[3] = {d->qvel[3], d->qvel[4], d->qvel[5]}, ang = h * norm3(w);
  if (ang > 0) { /* quat <- quat * exp(h*omega_body) */
    double n = norm3(w), s = sin(0.5 * ang), qr[4] = {cos(0.5 * ang), s * w[0] / n, s * w[1] / n, s * w[2] / n}, q2[4];
    quatMul(q2, d->qpos + 3, qr);
    memcpy(d->qpos + 3, q2, sizeof(q2));
  }
  quatNormalize(d->qpos + 3);
  for (int j = 0; j < QS_NJNT; j++) d->qpos[7 + j] += h * d->qvel[6 + j];
  d->time += h;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc));
}

/* ------------------------------------------------------------------ env side (quadruped_env.py) */
static void envFlags(OData* d) {
  /* feet_contact_state :836-847, _check_for_invalid_contacts :1232-1244, _check_out_of_terrain_bounds :1252-1256 */
  memset(d->contact_state, 0, sizeof(d->contact_state));
  d->invalid_contact = 0;
  d->invalid_body_mask = 0;
  for (int ci = 0; ci < d->ncon; ci++) {
    int b = d->con[ci].body, leg = (b >= 2 && (b - 2) % 3 == 2) ? (b - 2) / 3 : -1; /* calf body owns the foot geom (:1366) */
    if (leg >= 0) d->contact_state[leg] = 1;
    else { d->invalid_contact = 1; d->invalid_body_mask |= 1u << b; }
  }
  const double* L = d->m.terrain_limits;
  d->out_of_bounds = d->qpos[0] > L[0] || d->qpos[0] < L[1] || d->qpos[1] > L[2] || d->qpos[1] < L[3];
}

/* ALL_OBS pack in the order of SURVEY.md section 8(a); legs in model order FL,FR,RL,RR (= default legs_order :95) */
static void packObs(OData* d, double* o) {
  double R[9], q[4];
  memcpy(q, d->qpos + 3, sizeof(q));
  { double n = sqrt(q[0]*q[0]+q[1]*q[1]+q[2]*q[2]+q[3]*q[3]); for (int i = 0; i < 4; i++) q[i] /= n; } /* scipy from_quat normalises (:968) */
  quat2Mat(R, q);
  double roll = atan2(R[7], R[8]), pitch = -asin(fmax(-1.0, fmin(1.0, R[6]))), yaw = atan2(R[3], R[0]); /* as_euler('xyz') :987 */
  double cy = cos(yaw), sy = sin(yaw);
  double vH[3] = {d->command[0], d->command[1], d->command[2]}, yawrate = d->command[3];
  double vref[3] = {cy * vH[0] - sy * vH[1], sy * vH[0] + cy * vH[1], vH[2]}; /* heading_orientation_SO3 @ v_H :492-493 */
  double wref[3] = {0, 0, yawrate};
  const double *v = d->qvel, *wb = d->qvel + 3;
  double tmp[3], tmp2[3];
  int k = 0;
  for (int i = 0; i < 3; i++) o[k++] = d->qpos[i];                       /* base_pos */
  for (int i = 0; i < 3; i++) o[k++] = v[i];                             /* base_lin_vel */
  for (int i = 0; i < 3; i++) o[k++] = vref[i] - v[i];                   /* base_lin_vel_err */
  for (int i = 0; i < 3; i++) o[k++] = d->qacc[i];                       /* base_lin_acc :536 */
  mulMatVec3(tmp, R, wb);
  for (int i = 0; i < 3; i++) o[k++] = tmp[i];                           /* base_ang_vel :529 */
  for (int i = 0; i < 3; i++) o[k++] = wref[i] - tmp[i];                 /* base_ang_vel_err */
  o[k++] = roll; o[k++] = pitch; o[k++] = yaw;                           /* base_ori_euler_xyz */
  for (int i = 0; i < 4; i++) o[k++] = d->qpos[3 + i];                   /* base_ori_quat_wxyz :1180 */
  for (int i = 0; i < 9; i++) o[k++] = R[i];                             /* base_ori_SO3 */
  { double g[3] = {0, 0, -1}; mulMatTVec3(tmp, R, g); for (int i = 0; i < 3; i++) o[k++] = tmp[i]; } /* gravity_vector:base */
  mulMatTVec3(tmp, R, v);
  for (int i = 0; i < 3; i++) o[k++] = tmp[i];                           /* base_lin_vel:base */
  mulMatTVec3(tmp2, R, vref);
  for (int i = 0; i < 3; i++) o[k++] = tmp2[i] - tmp[i];                 /* base_lin_vel_err:base */
  mulMatTVec3(tmp, R, d->qacc);
  for (int i = 0; i < 3; i++) o[k++] = tmp[i];                           /* base_lin_acc:base */
  for (int i = 0; i < 3; i++) o[k++] = wb[i];                            /* base_ang_vel:base */
  mulMatTVec3(tmp, R, wref);
  for (int i = 0; i < 3; i++) o[k++] = tmp[i] - wb[i];                   /* base_ang_vel_err:base */
  for (int i = 0; i < NQ; i++) o[k++] = d->qpos[i];
  for (int i = 0; i < NV; i++) o[k++] = d->qvel[i];
  for (int i = 0; i < NU; i++) o[k++] = d->ctrl[i];                      /* tau_ctrl_setpoint: unclamped :1005 */
  for (int i = 0; i < 12; i++) o[k++] = d->qpos[7 + i];
  for (int i = 0; i < 12; i++) o[k++] = d->qvel[6 + i];
  { /* kinetic_energy, work: intended formulas (:941, :956-957), M and qacc from the forward pass */
    double ke = 0, wk = 0;
    for (int i = 0; i < NV; i++) { double mv = 0, ma = 0; for (int j = 0; j < NV; j++) { mv += d->M[i][j] * d->qvel[j]; ma += d->M[i][j] * d->qacc[j]; } ke += 0.5 * d->qvel[i] * mv; wk += ma * d->qvel[i]; }
    o[k++] = ke; o[k++] = wk;
  }
  double fpos[4][3], fvel[4][3], frel[4][3];
  for (int l = 0; l < 4; l++) {
    int g = d->m.foot_geom[l], calf = 4 + 3 * l;
    memcpy(fpos[l], d->geom_xpos[g], sizeof(fpos[l]));
    double jp[3][NV];
    jacPoint(d, jp, NULL, fpos[l], calf);                                /* mj_jac :728-735 */
    for (int i = 0; i < 3; i++) { double s = 0; for (int j = 0; j < NV; j++) s += jp[i][j] * d->qvel[j]; fvel[l][i] = s; }
    double r[3] = {fpos[l][0] - d->qpos[0], fpos[l][1] - d->qpos[1], fpos[l][2] - d->qpos[2]}, c[3];
    cross3(c, wb, r);                                                    /* body-frame omega used as world :659,:669 */
    for (int i = 0; i < 3; i++) frel[l][i] = fvel[l][i] - v[i] - c[i];
  }
  for (int l = 0; l < 4; l++) for (int i = 0; i < 3; i++) o[k++] = fpos[l][i];            /* feet_pos */
  for (int l = 0; l < 4; l++) { for (int i = 0; i < 3; i++) tmp[i] = fpos[l][i] - d->qpos[i]; mulMatTVec3(tmp2, R, tmp); for (int i = 0; i < 3; i++) o[k++] = tmp2[i]; } /* feet_pos:base :620 */
  for (int l = 0; l < 4; l++) for (int i = 0; i < 3; i++) o[k++] = fvel[l][i];
  for (int l = 0; l < 4; l++) for (int i = 0; i < 3; i++) o[k++] = frel[l][i];
  for (int l = 0; l < 4; l++) { mulMatTVec3(tmp, R, fvel[l]); for (int i = 0; i < 3; i++) o[k++] = tmp[i]; }
  for (int l = 0; l < 4; l++) { mulMatTVec3(tmp, R, frel[l]); for (int i = 0; i < 3; i++) o[k++] = tmp[i]; }
  for (int l = 0; l < 4; l++) o[k++] = d->contact_state[l];
  double cf[4][3];
  memset(cf, 0, sizeof(cf));
  for (int ci = 0; ci < d->ncon; ci++) { /* R_c^T f[:3] summed per leg :849-855 */
    OContact* c = &d->con[ci];
    int b = c->body, leg = (b >= 2 && (b - 2) % 3 == 2) ? (b - 2) / 3 : -1;
    if (leg < 0) continue;
    for (int i = 0; i < 3; i++) cf[leg][i] += c->frame[i] * c->force[0] + c->frame[3 + i] * c->force[1] + c->frame[6 + i] * c->force[2];
  }
  for (int l = 0; l < 4; l++) for (int i = 0; i < 3; i++) o[k++] = cf[l][i];
  for (int l = 0; l < 4; l++) { mulMatTVec3(tmp, R, cf[l]); for (int i = 0; i < 3; i++) o[k++] = tmp[i]; }
}

/* ------------------------------------------------------------------ exported API (ctypes) */
void* orc_create(const QsModel* m) {
  OData* d = (OData*)calloc(1, sizeof(OData));
  d->m = *m;
  if (m->nvert > 0) { d->vert = (double*)malloc(sizeof(double) * 3 * m->nvert); memcpy(d->vert, m->vert, sizeof(double) * 3 * m->nvert); }
  if (m->hf_data && m->hf_nrow > 0) { size_t n = (size_t)m->hf_nrow * m->hf_ncol; d->hf = (float*)malloc(sizeof(float) * n); memcpy(d->hf, m->hf_data, sizeof(float) * n); }
  d->m.vert = d->vert; d->m.hf_data = d->hf;
  memcpy(d->qpos, m->key_qpos, sizeof(d->qpos));
  d->mu_floor = d->mu_feet = -1;
  d->ctx = calloc(1, sizeof(Ctx));
  return d;
}
void orc_destroy(void* h) { OData* d = (OData*)h; free(d->vert); free(d->hf); free(d->ctx); free(d); }
int orc_model_sizeof(void) { return (int)sizeof(QsModel); }

void orc_set_state(void* h, const double* qpos, const double* qvel, const double* warm) {
  OData* d = (OData*)h;
  if (qpos) memcpy(d->qpos, qpos, sizeof(d->qpos));
  if (qvel) memcpy(d->qvel, qvel, sizeof(d->qvel));
  if (warm) memcpy(d->qacc_warmstart, warm, sizeof(d->qacc_warmstart));
}
void orc_get_state(void* h, double* qpos, double* qvel, double* qacc, double* warm) {
  OData* d = (OData*)h;
  if (qpos) memcpy(qpos, d->qpos, sizeof(d->qpos));
  if (qvel) memcpy(qvel, d->qvel, sizeof(d->qvel));
  if (qacc) memcpy(qacc, d->qacc, sizeof(d->qacc));
  if (warm) memcpy(warm, d->qacc_warmstart, sizeof(d->qacc_warmstart));
}
void orc_set_env(void* h, double mu_floor, double mu_feet, const double* command, const double* qfrc_applied6) {
  OData* d = (OData*)h;
  d->mu_floor = mu_floor; d->mu_feet = mu_feet;
  if (command) memcpy(d->command, command, sizeof(d->command));
  if (qfrc_applied6) memcpy(d->qfrc_applied, qfrc_applied6, 6 * sizeof(double));
}
void orc_set_time(void* h, double t) { ((OData*)h)->time = t; }
double orc_get_time(void* h) { return ((OData*)h)->time; }

void orc_forward(void* h, const double* ctrl) {
  OData* d = (OData*)h;
  if (ctrl) memcpy(d->ctrl, ctrl, sizeof(d->ctrl));
  forward(d);
  envFlags(d);
}

/* QuadrupedEnv.step :270-288: returns terminated flag; obs may be NULL. obs has 227 (+6 truth IMU when has_imu) doubles */
int orc_step(void* h, const double* ctrl, double* obs) {
  OData* d = (OData*)h;
  memcpy(d->ctrl, ctrl, sizeof(d->ctrl));
  forward(d);
  euler(d);
  envFlags(d);
  if (obs) {
    packObs(d, obs);
    if (d->m.has_imu) { memcpy(obs + QS_NOBS_BASE, d->sensor_acc, sizeof(d->sensor_acc)); memcpy(obs + QS_NOBS_BASE + 3, d->sensor_gyro, sizeof(d->sensor_gyro)); }
  }
  /* libqstep rule (no counterpart in the env; the engine itself warns and resets its data on a bad state): a non-finite state ends the episode */
  int bad = 0;
  for (int i = 0; i < QS_NQ; i++) bad |= !isfinite(d->qpos[i]);
  for (int i = 0; i < QS_NV; i++) bad |= !isfinite(d->qvel[i]);
  return d->invalid_contact || d->out_of_bounds || bad;
}

/* reset lift loop, quadruped_env.py:376-388: returns number of lifts, -1 if contact could not be cleared */
int orc_lift(void* h) {
  OData* d = (OData*)h;
  for (int c = 0; c <= 100; c++) {
    kinematics(d);
    collision(d);
    envFlags(d);
    double maxpen = 0;
    int any = 0;
    for (int ci = 0; ci < d->ncon; ci++) {
      int b = d->con[ci].body;
      if (b >= 2 && (b - 2) % 3 == 2) { any = 1; if (fabs(d->con[ci].dist) > maxpen) maxpen = fabs(d->con[ci].dist); }
    }
    if (!any) return c;
    if (c == 100) break;
    d->qpos[2] += maxpen * 1.1;
  }
  return -1;
}

enum { F_M = 0, F_BIAS = 1, F_PASSIVE = 2, F_FEET_JACP = 3, F_FEET_POS = 4, F_COM = 5, F_CONTACTS = 6, F_SMOOTH = 7, F_CONSTRAINT = 8,
       F_XPOS = 9, F_IMU = 10, F_QACC_SMOOTH = 11, F_EFC = 12, F_FLAGS = 13, F_FEET_JACR = 14, F_FEET_JACP_DOT = 15, F_FEET_JACR_DOT = 16,
       F_EFC_FULL = 17 };
#define EFC_FULL_STRIDE 34
int orc_get(void* h, int field, double* dst) {
  OData* d = (OData*)h;
  switch (field) {
    case F_M: memcpy(dst, d->M, sizeof(d->M)); return NV * NV;
    case F_BIAS: memcpy(dst, d->qfrc_bias, sizeof(d->qfrc_bias)); return NV;
    case F_PASSIVE: memcpy(dst, d->qfrc_passive, sizeof(d->qfrc_passive)); return NV;
    case F_FEET_JACP:
      for (int l = 0; l < 4; l++) { double jp[3][NV]; jacPoint(d, jp, NULL, d->geom_xpos[d->m.foot_geom[l]], 4 + 3 * l); memcpy(dst + l * 3 * NV, jp, sizeof(jp)); }
      return 4 * 3 * NV;
    case F_FEET_JACR: case F_FEET_JACP_DOT: case F_FEET_JACR_DOT:
      for (int l = 0; l < 4; l++) {
        double jp[3][NV], jr[3][NV];
        if (field == F_FEET_JACR) jacPoint(d, jp, jr, d->geom_xpos[d->m.foot_geom[l]], 4 + 3 * l);
        else jacDotPoint(d, jp, jr, d->geom_xpos[d->m.foot_geom[l]], 4 + 3 * l);
        memcpy(dst + l * 3 * NV, field == F_FEET_JACP_DOT ? jp : jr, sizeof(jp));
      }
      return 4 * 3 * NV;
    case F_FEET_POS: for (int l = 0; l < 4; l++) memcpy(dst + 3 * l, d->geom_xpos[d->m.foot_geom[l]], 3 * sizeof(double)); return 12;
    case F_COM: memcpy(dst, d->com, sizeof(d->com)); return 3;
    case F_CONTACTS:
      for (int ci = 0; ci < d->ncon; ci++) {
        OContact* c = &d->con[ci];
        double* o = dst + QS_CONTACT_STRIDE * ci;
        o[0] = c->dist; memcpy(o + 1, c->pos, 3 * sizeof(double)); memcpy(o + 4, c->frame, 9 * sizeof(double));
        memcpy(o + 13, c->force, 3 * sizeof(double)); o[16] = c->geom; o[17] = c->body; o[18] = c->friction[0]; o[19] = c->dim;
      }
      return d->ncon;
    case F_SMOOTH: memcpy(dst, d->qfrc_smooth, sizeof(d->qfrc_smooth)); return NV;
    case F_CONSTRAINT: memcpy(dst, d->qfrc_constraint, sizeof(d->qfrc_constraint)); return NV;
    case F_XPOS: memcpy(dst, d->xpos[1], 13 * 3 * sizeof(double)); return 39;
    case F_IMU: memcpy(dst, d->sensor_acc, sizeof(d->sensor_acc)); memcpy(dst + 3, d->sensor_gyro, sizeof(d->sensor_gyro)); return 6;
    case F_QACC_SMOOTH: memcpy(dst, d->qacc_smooth, sizeof(d->qacc_smooth)); return NV;
    case F_EFC: /* per row: type, D, R, aref, force, state */
      for (int r = 0; r < d->nefc; r++) { double* o = dst + 6 * r; o[0] = d->efc_type[r]; o[1] = d->efc_D[r]; o[2] = d->efc_R[r]; o[3] = d->efc_aref[r]; o[4] = d->efc_force[r]; o[5] = d->efc_state[r]; }
      return d->nefc;
    case F_EFC_FULL: /* the constraint problem handed to the solver, one row each: type, id, D, R, aref, force, state, floss, contact mu,
                        friction[5], contact dim, pad, J[18] -- for the independent-optimiser pin in tests/test_solver_pin.py */
      for (int r = 0; r < d->nefc; r++) {
        double* o = dst + EFC_FULL_STRIDE * r;
        int type = d->efc_type[r], id = d->efc_id[r];
        memset(o, 0, EFC_FULL_STRIDE * sizeof(double));
        o[0] = type; o[1] = id; o[2] = d->efc_D[r]; o[3] = d->efc_R[r]; o[4] = d->efc_aref[r]; o[5] = d->efc_force[r]; o[6] = d->efc_state[r];
        o[7] = d->efc_floss[r];
        if (type >= T_CONTACT_FRICTIONLESS) { const OContact* c = &d->con[id]; o[8] = c->mu; memcpy(o + 9, c->friction, 5 * sizeof(double)); o[14] = c->dim; }
        memcpy(o + 16, d->efc_J[r], NV * sizeof(double));
      }
      return d->nefc;
    case F_FLAGS:
      for (int l = 0; l < 4; l++) dst[l] = d->contact_state[l];
      dst[4] = d->invalid_contact; dst[5] = d->out_of_bounds; dst[6] = d->ncon; dst[7] = d->nefc; dst[8] = d->solver_iter; dst[9] = d->overflow; dst[10] = d->invalid_body_mask;
      return 11;
    default: return -1;
  }
}

/* HeightMap.create_sensor_matrix (sensors/heightmap.py:106-169): rows x cols points [rows][cols][3] around `center` with heading yaw */
void orc_heightmap(void* h, const double* center, double yaw, int rows, int cols, double dx, double dy, double* out) {
  OData* d = (OData*)h;
  double c_rows = rows % 2 == 0 ? rows / 2.0 : (rows - 1) / 2.0, add_r = rows % 2 == 0 ? -dx / 2.0 : 0.0;
  double c_cols = cols % 2 == 0 ? cols / 2.0 : (cols - 1) / 2.0, add_c = cols % 2 == 0 ? -dy / 2.0 : 0.0;
  double cy = cos(yaw), sy = sin(yaw);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      double ox = dx * (c_rows - i) + add_r, oy = dy * (c_cols - j) + add_c;
      double org[3] = {center[0] + cy * ox - sy * oy, center[1] + sy * ox + cy * oy, center[2] + 0.6 - 0.07};
      double t = rayDown(d, org);
      double* o = out + 3 * (i * cols + j);
      o[0] = org[0]; o[1] = org[1]; o[2] = org[2] - t; /* a miss (t = -1) lands 1 m above the ray origin (:103) */
    }
}

/* K steps with a fixed ctrl table (K x 12) for CPU-baseline timing; returns number of terminated steps */
int orc_rollout(void* h, const double* ctrl, int K, double* obs_last) {
  int term = 0;
  double obs[QS_NOBS_BASE + 6];
  for (int k = 0; k < K; k++) term += orc_step(h, ctrl + NU * k, obs);
  if (obs_last) memcpy(obs_last, obs, sizeof(obs));
  return term;
}

/* Rollout with auto-reset for CPU-baseline timing: on termination the env is re-initialised from the next entry of a
 * table of pre-lifted start states (nreset x (19 qpos + 18 qvel)), mirroring the GPU benchmark loop. Returns #resets. */
int orc_rollout_autoreset(void* h, const double* ctrl, int K, const double* reset_states, int nreset, int* cursor) {
  OData* d = (OData*)h;
  int resets = 0;
  double obs[QS_NOBS_BASE + 6];
  for (int k = 0; k < K; k++) {
    if (orc_step(h, ctrl + NU * k, obs)) {
      const double* s = reset_states + (size_t)((*cursor) % nreset) * (NQ + NV);
      (*cursor)++;
      memcpy(d->qpos, s, sizeof(d->qpos));
      memcpy(d->qvel, s + NQ, sizeof(d->qvel));
      memset(d->qacc_warmstart, 0, sizeof(d->qacc_warmstart));
      memset(d->ctrl, 0, sizeof(d->ctrl));
      d->time = 0;
      forward(d);
      euler(d); /* reset() performs one full step, quadruped_env.py:397 */
      resets++;
    }
  }
  return resets;
}
