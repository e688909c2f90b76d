/*
 * qstep.h -- C-ABI of the B200-native batched quadruped step ("libqstep").
 *
 * This is the drop-in boundary for the hot path of iit-DLSLab/gym-quadruped:
 *   QuadrupedEnv.step   gym_quadruped/quadruped_env.py:251-307
 *   QuadrupedEnv.reset  gym_quadruped/quadruped_env.py:309-406
 *   QuadrupedEnv._get_obs / _check_for_invalid_contacts / _check_out_of_terrain_bounds
 *                       gym_quadruped/quadruped_env.py:1146-1257
 *   mujoco.mj_step      (third-party engine called at quadruped_env.py:271,397)
 *   HeightMap.update_height_map  gym_quadruped/sensors/heightmap.py:66-169
 *   IMU.step            gym_quadruped/sensors/imu.py:102-139
 *
 * The reference has no FFI of its own for this path (its only native boundary is the
 * `mujoco` pybind11 module); the entry points below are what a maintainer binds with
 * ctypes in place of the MjModel/MjData calls -- see INTEGRATION.md.
 *
 * Conventions
 *  - plain C structs, pointers and sizes; no torch / C++ types in any signature.
 *  - every `dev` pointer is a CUDA device pointer owned by the caller (PyTorch);
 *    the library owns only its handle and the constant tables copied at create time.
 *  - all calls enqueue work on the given stream and do NOT synchronise, except the
 *    `*_host` variants, which copy host<->device and synchronise the stream.
 *  - return 0 on success, non-zero error code otherwise; text via qs_last_error().
 *  - per-env numerical failure is reported in-band (QsBuffers.status), never as an error.
 */
#ifndef QSTEP_H_
#define QSTEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_ABI_VERSION 5

/* ---- fixed topology of every robot in robot_cfgs.py: ----
 * world -> base(free joint) -> 4 x {hip, thigh, calf} (one hinge each)           */
#define QS_NBODY 14 /* world + base + 12 links                                     */
#define QS_NJNT 12  /* hinge joints; joint j lives on body j+2, dof 6+j, qpos 7+j  */
#define QS_NQ 19
#define QS_NV 18
#define QS_NU 12
#define QS_NLEG 4
#define QS_MAXGEOM 48  /* robot collision geoms (go2: 31, go1: 42)                  */
#define QS_MAXBOX 128  /* static boxes of the random_boxes scene (terrain.py:145)   */
#define QS_NOBS_BASE 227 /* ALL_OBS, quadruped_env.py:35-81                          */
#define QS_NOBS_IMU 18   /* sensors/imu.py:16-17                                     */

/* geom type ids follow the engine's ordering (contact geom1/geom2 are sorted by it) */
enum { QS_GEOM_PLANE = 0, QS_GEOM_HFIELD = 1, QS_GEOM_SPHERE = 2, QS_GEOM_CAPSULE = 3,
       QS_GEOM_ELLIPSOID = 4, QS_GEOM_CYLINDER = 5, QS_GEOM_BOX = 6, QS_GEOM_MESH = 7 };

enum { QS_CONE_PYRAMIDAL = 0, QS_CONE_ELLIPTIC = 1 };
enum { QS_TERRAIN_FLAT = 0, QS_TERRAIN_HFIELD = 1, QS_TERRAIN_BOXES = 2 };

/* contact-parameter block shared by every geom (robot, floor, hfield, boxes) */
typedef struct QsGeomParams {
  double friction[3]; /* slide, spin, roll */
  double solref[2];
  double solimp[5];
  double solmix;
  double margin;
  double gap;
  int32_t condim;
  int32_t priority;
} QsGeomParams;

/* Compiled robot + scene ("MjModel" subset). Produced by gym_quadruped_b200.compiler from
 * the reference's MJCF (robot_model/<robot>/<robot>.xml + utils/mujoco/assets/scene_flat.xml). */
typedef struct QsModel {
  int32_t abi_version;
  /* <option> */
  double timestep;
  double gravity[3];
  double impratio;
  double tolerance;
  double ls_tolerance;
  double meaninertia; /* stat.meaninertia at compile-time qpos0 */
  int32_t cone;
  int32_t iterations;
  int32_t ls_iterations;
  int32_t noslip_pad;

  /* bodies; index 0 = world */
  int32_t body_parent[QS_NBODY];
  double body_pos[QS_NBODY][3];
  double body_quat[QS_NBODY][4];
  double body_ipos[QS_NBODY][3];
  double body_iquat[QS_NBODY][4];
  double body_mass[QS_NBODY];
  double body_inertia[QS_NBODY][3];
  double body_invweight0[QS_NBODY][2];

  /* hinge joints */
  double jnt_pos[QS_NJNT][3];
  double jnt_axis[QS_NJNT][3];
  double jnt_range[QS_NJNT][2];
  double jnt_solref[QS_NJNT][2];
  double jnt_solimp[QS_NJNT][5];
  double jnt_margin[QS_NJNT];
  int32_t jnt_limited[QS_NJNT];

  double qpos0[QS_NQ];     /* run-time reference pose (after quadruped_env.py:171-173) */
  double key_qpos[QS_NQ];  /* keyframe 0 ("home")                                      */

  double dof_damping[QS_NV];
  double dof_armature[QS_NV];
  double dof_frictionloss[QS_NV];
  double dof_invweight0[QS_NV];
  double dof_solref[QS_NV][2];
  double dof_solimp[QS_NV][5];

  /* motors: actuator a drives hinge a (gear 1) */
  double act_ctrlrange[QS_NU][2];
  double act_forcerange[QS_NU][2];
  int32_t act_ctrllimited[QS_NU];
  int32_t act_forcelimited[QS_NU];

  /* robot collision geoms (contype/conaffinity != 0 only), in model order */
  int32_t ngeom;
  int32_t geom_type[QS_MAXGEOM];
  int32_t geom_body[QS_MAXGEOM];
  int32_t geom_foot_leg[QS_MAXGEOM]; /* leg index (model order FL,FR,RL,RR) if this is a foot geom, else -1 */
  int32_t geom_vertadr[QS_MAXGEOM];  /* mesh: first hull vertex in vert[]  */
  int32_t geom_vertnum[QS_MAXGEOM];
  double geom_pos[QS_MAXGEOM][3];    /* in body frame */
  double geom_quat[QS_MAXGEOM][4];
  double geom_size[QS_MAXGEOM][3];
  double geom_bcenter[QS_MAXGEOM][3]; /* bounding sphere (body frame) */
  double geom_rbound[QS_MAXGEOM];
  QsGeomParams geom_par[QS_MAXGEOM];
  int32_t foot_geom[QS_NLEG];        /* geom index of FL,FR,RL,RR foot spheres */

  /* convex-hull vertices of mesh geoms, expressed in the owning BODY frame */
  int32_t nvert;
  int32_t pad0;
  const double* vert; /* [nvert][3], host pointer, copied at create time */

  /* scene */
  int32_t terrain_type;
  int32_t nbox;
  QsGeomParams floor_par;     /* plane z=0, scene_flat.xml:32 */
  double terrain_limits[4];   /* (x_max, x_min, y_max, y_min), terrain.py:118,237,359 */
  /* hfield (perlin scene, terrain.py:25-119) */
  int32_t hf_nrow, hf_ncol;
  double hf_size[4];          /* half-x, half-y, z-scale, base thickness */
  double hf_pos[3];
  const float* hf_data;       /* [nrow][ncol] in [0,1], host pointer */
  QsGeomParams hf_par;
  /* static boxes (random_boxes scene, terrain.py:145-238) */
  double box_pos[QS_MAXBOX][3];
  double box_quat[QS_MAXBOX][4];
  double box_half[QS_MAXBOX][3];
  double box_friction[QS_MAXBOX][3]; /* per-box slide / spin / roll friction (scene_slippery.xml:39-40) */
  QsGeomParams box_par;              /* everything else is shared by the boxes of a scene (priority 2 in `slippery`) */

  /* IMU site (sensors/imu.py): accelerometer + gyro attached to a site on the base body */
  int32_t has_imu;
  int32_t pad1;
  double imu_pos[3];
  double imu_quat[4];
} QsModel;

typedef struct QsConfig {
  int32_t num_envs;
  int32_t device;          /* CUDA device ordinal */
  int32_t precision;       /* 0 = fp32 state/arithmetic (product), 1 = fp64 arithmetic (parity / debug build of the same kernel) */
  int32_t use_imu;         /* append 18 IMU columns to the obs row */
  int32_t hm_rows, hm_cols; /* height-map grid appended to the obs row (0 = none) */
  double hm_dx, hm_dy;     /* grid spacing, sensors/heightmap.py:22-27 */
  double imu_accel_noise, imu_gyro_noise, imu_accel_bias_rate, imu_gyro_bias_rate; /* imu.py:36-41 */
  uint64_t seed;
  int32_t env_id_offset;   /* global id of local env 0 (multi-GPU sharding; keys the counter RNG) */
  int32_t solver_max_iter; /* <=0: library default */
  int32_t pipeline;        /* 1: consecutive qs_step / qs_step_autoreset launches may overlap on the device (programmatic dependent
                              launch + per-env finish-order queues; results are bit-identical to the serialized order).  Contract: the
                              buffers a step reads (ctrl, bound buffers edited by the caller) must be complete when the PREVIOUS step
                              launch of this handle was enqueued, or be written by the host.  0 (default): plain stream order. */
  int32_t pad0;
} QsConfig;

/* Device buffers owned by the caller (torch tensors), bound once after qs_create.
 * Layout: row-major [num_envs, width]; fp32 unless noted. */
typedef struct QsBuffers {
  float* qpos;          /* [N,19]  (base xyz here is the fp32 image of base_pos64)            */
  float* qvel;          /* [N,18]                                                              */
  float* qacc;          /* [N,18]  solver acceleration of the last forward pass               */
  float* qacc_warmstart;/* [N,18]                                                              */
  double* base_pos64;   /* [N,3]   fp64 master copy of qpos[0:3] (flat resets span +-1e4 m)   */
  float* qfrc_applied;  /* [N,6]   external wrench on the base dofs, quadruped_env.py:305     */
  float* command;       /* [N,4]   (v_H.x, v_H.y, v_H.z, yaw_rate) quadruped_env.py:1046-1072 */
  float* friction;      /* [N,2]   (floor mu, feet mu); <0 = model value, quadruped_env.py:1277 */
  float* sim_time;      /* [N]                                                                 */
  int32_t* step_count;  /* [N]                                                                 */
  float* imu_bias;      /* [N,6]   accel bias, gyro bias (random walk, imu.py:123,136)        */
  uint8_t* status;      /* [N]     bit0 = non-finite state, bit1 = contact buffer overflow, bit2 = solver hit max iterations */
  int32_t* ncon;        /* [N]     number of active contacts in the last forward pass         */
  int32_t* solver_iter; /* [N]                                                                 */
  uint8_t* invalid_body_mask; /* [N,2] bitmask (little endian u16) of robot bodies with a world contact that is not a foot/calf body */
  /* in-episode schedules, quadruped_env.py:293-305 (see QsSchedule) */
  int32_t* cmd_count;   /* [N]   steps since the velocity command was last drawn ('+reset' command types, :293-296)           */
  int32_t* cmd_limit;   /* [N]   randint(1000, 3000) drawn together with the command (:1068-1070)                            */
  int32_t* ext_count;   /* [N]   steps since the external wrench was last drawn (:299-302)                                   */
  int32_t* ext_limit;   /* [N]                                                                                               */
  float* ext_wrench;    /* [N,6] current disturbance (x, y, z, roll, pitch, yaw); copied to qfrc_applied after every step (:305) */
} QsBuffers;

/* In-episode schedules run inside the step kernel (no host round trip): resampling of the velocity command for '+reset' command
 * types (quadruped_env.py:293-296, _sample_ref_vel :1046-1072) and of the external base wrench for
 * external_disturbances_kwargs['type'] == 'reset' (:299-305, _sample_external_disturbances :1074-1139). */
typedef struct QsSchedule {
  int32_t command_mode;     /* as QsResetOptions.command_mode; bit3 enables the in-episode command resampling            */
  int32_t ext_enabled;      /* 1: disturbance schedule on (kwargs given and type == 'reset')                             */
  double lin_vel_range[2];
  double ang_vel_range[2];
  double ext_lo[6], ext_hi[6]; /* per component U(lo, hi); lo == hi: fixed value (a one-element range or an absent key = 0) */
} QsSchedule;

/* options of a random reset, quadruped_env.py:346-373 */
typedef struct QsResetOptions {
  double angle_sweep;  /* default 20 deg */
  double vel_sweep;    /* 0.5            */
  double roll_sweep;   /* default 10 deg */
  double pitch_sweep;
  double hip_height;   /* robot_cfgs.py  */
  double lin_vel_range[2];
  double ang_vel_range[2];
  double friction_range[2];
  int32_t command_mode; /* bit0 forward, bit1 random heading, bit2 rotate, bit3 resample-on-schedule */
  int32_t randomize;    /* reset(random=...) */
} QsResetOptions;

typedef struct QsHandle_ QsHandle;

/* library / build info */
int qs_abi_version(void);
int qs_model_sizeof(void);
int qs_config_sizeof(void);
int qs_buffers_sizeof(void);
int qs_obs_dim(const QsConfig* cfg);

/* lifetime (replaces MjModel.from_xml_path + MjData, quadruped_env.py:170,178) */
int qs_create(const QsModel* model, const QsConfig* cfg, QsHandle** out);
void qs_destroy(QsHandle* h);
const char* qs_last_error(QsHandle* h);
int qs_bind(QsHandle* h, const QsBuffers* dev_buffers);

/* Install the in-episode schedules and draw the initial wrench / limits of every env (the reference samples them in __init__,
 * quadruped_env.py:240-242).  NULL switches both schedules off. */
int qs_set_schedule(QsHandle* h, const QsSchedule* sched, void* cuda_stream);

/* `np.random.seed(seed)` of reset(seed=...) (quadruped_env.py:337-338) for the counter-based generator: replaces the key and
 * restarts the per-env draw counters (episode, IMU tick, schedule epochs); no buffer is touched. */
int qs_set_seed(QsHandle* h, uint64_t seed, void* cuda_stream);

/* name of the compiled kernel variant that qs_step launches for this handle (diagnostics / tests) */
const char* qs_step_variant(QsHandle* h);

/* mj_step + sensors + _get_obs + termination, quadruped_env.py:270-288.
 * ctrl [N,12]; obs [N,D]; reward [N]; terminated/truncated [N]. */
int qs_step(QsHandle* h, const float* dev_ctrl, float* dev_obs, float* dev_reward,
            uint8_t* dev_terminated, uint8_t* dev_truncated, void* cuda_stream);

/* qs_step followed, inside the same kernel launch, by a random reset (as qs_reset with the given options) of every env
 * that just terminated.  For those envs `terminated` stays 1 and obs / state hold the post-reset values (the usual
 * vectorised-rollout convention); the reference leaves this loop to the caller (quadruped_env.py:1419-1421). */
int qs_step_autoreset(QsHandle* h, const float* dev_ctrl, const QsResetOptions* opt, float* dev_obs, float* dev_reward,
                      uint8_t* dev_terminated, uint8_t* dev_truncated, void* cuda_stream);

/* K consecutive steps in one call: step i uses dev_ctrl[i] ([K,N,12]) and writes its observation rows to
 * dev_obs + i * obs_step_stride floats (stride 0: every step overwrites the same [N,D] tensor; stride N*D: an observation ring),
 * its reward / flags to row i of [K,N] arrays (NULL allowed).  Bit-identical to K calls of qs_step / qs_step_autoreset
 * (auto_reset NULL / non-NULL).  It is still one kernel launch per step, but because the library issues the launches back to
 * back it lets each overlap its predecessor on the device (see QsConfig.pipeline) without any contract on the caller: the first
 * launch of the sequence is ordered after everything enqueued before the call.  For open-loop action segments (replay, MPC
 * roll-outs, the benchmark). */
int qs_step_k(QsHandle* h, int k, const float* dev_ctrl, const QsResetOptions* auto_reset, float* dev_obs, size_t obs_step_stride,
              float* dev_reward, uint8_t* dev_terminated, uint8_t* dev_truncated, void* cuda_stream);

/* same call with HOST buffers: H2D(ctrl) -> step -> D2H(obs, reward, flags), stream-synchronised.
 * auto_reset: NULL = plain qs_step, else qs_step_autoreset with these options. */
int qs_step_host(QsHandle* h, const float* host_ctrl, const QsResetOptions* auto_reset, float* host_obs, float* host_reward,
                 uint8_t* host_terminated, uint8_t* host_truncated, void* cuda_stream);

/* qs_step_host with padded observation rows: row i starts at host_obs + i * obs_row_stride floats (>= D; 0 = D).  With a pinned
 * buffer whose rows start on 128-byte boundaries (e.g. stride 256 floats for D = 227) the zero-copy row writes cross PCIe as
 * whole lines instead of two partial lines per row. */
int qs_step_host_strided(QsHandle* h, const float* host_ctrl, const QsResetOptions* auto_reset, float* host_obs, size_t obs_row_stride,
                         float* host_reward, uint8_t* host_terminated, uint8_t* host_truncated, void* cuda_stream);

/* reset, quadruped_env.py:309-406. env_mask [N] (NULL = all). If dev_qpos/dev_qvel are non-NULL
 * ([N,19]/[N,18]) they are taken verbatim (the `else` branch, :389-391), otherwise keyframe + noise +
 * lift-until-no-foot-contact. Always followed by one full step with zero ctrl (:397) and an obs pack. */
int qs_reset(QsHandle* h, const uint8_t* dev_env_mask, const float* dev_qpos, const float* dev_qvel,
             const QsResetOptions* opt, float* dev_obs, void* cuda_stream);

/* auto-reset of every env whose dev_terminated flag is set (batched-rollout convenience; the reference
 * leaves this to the caller: quadruped_env.py:1419-1421). */
int qs_reset_done(QsHandle* h, const uint8_t* dev_terminated, const QsResetOptions* opt, float* dev_obs,
                  void* cuda_stream);

/* forward pass only (mj_forward, quadruped_env.py:1321), filling the accessor tables below. */
int qs_forward(QsHandle* h, void* cuda_stream);

/* accessor parity (mj_fullM :557, qfrc_bias :899, mj_jac :728, contacts :836, subtree_com :925) */
enum { QS_FIELD_MASS_MATRIX = 0, /* [N,18,18] */
       QS_FIELD_QFRC_BIAS = 1,   /* [N,18]    */
       QS_FIELD_QFRC_PASSIVE = 2,/* [N,18]    */
       QS_FIELD_FEET_JACP = 3,   /* [N,4,3,18] model leg order */
       QS_FIELD_FEET_POS = 4,    /* [N,4,3]   */
       QS_FIELD_COM = 5,         /* [N,3]     */
       QS_FIELD_CONTACTS = 6,    /* [N,QS_CONTACT_STRIDE * max_contacts] */
       QS_FIELD_QFRC_SMOOTH = 7, /* [N,18] */
       QS_FIELD_QFRC_CONSTRAINT = 8, /* [N,18] */
       QS_FIELD_XPOS = 9,        /* [N,13,3] body positions (bodies 1..13) */
       QS_FIELD_SENSOR_IMU = 10, /* [N,6] noiseless accelerometer + gyro */
       QS_FIELD_FEET_JACR = 11,  /* [N,4,3,18] rotational Jacobian of each calf body (mj_jac jacr, quadruped_env.py:728-735) */
       QS_FIELD_FEET_JACP_DOT = 12, /* [N,4,3,18] d/dt of QS_FIELD_FEET_JACP (mj_jacDot jacp, quadruped_env.py:785-792) */
       QS_FIELD_FEET_JACR_DOT = 13  /* [N,4,3,18] d/dt of QS_FIELD_FEET_JACR (mj_jacDot jacr) */ };
int qs_get(QsHandle* h, int field, float* dev_dst, void* cuda_stream);
int qs_max_contacts(QsHandle* h);
#define QS_CONTACT_STRIDE 20 /* dist, pos[3], frame[9], force[3], geom, body, mu, pad */

/* HeightMap.update_height_map, sensors/heightmap.py:106-169; out [N,rows,cols,3] */
int qs_raycast_heightmap(QsHandle* h, int rows, int cols, double dx, double dy, float* dev_out,
                         void* cuda_stream);

/* ---- peer-to-peer gather of the observation rows, fused into the step kernel (the one collective of the
 * path; one process per GPU, up to 8 GPUs of one node).  Every rank allocates its [2][world * N, D (row stride qs_gather_row_stride)] gathered tensors with
 * qs_gather_create and publishes the returned 64-byte CUDA IPC handle; after qs_gather_connect (all handles, rank-major) every
 * qs_step / qs_step_autoreset also stores each finished observation row into the gathered tensor of EVERY rank through
 * peer-mapped memory (the `dev_obs` argument is then ignored: the own rows live in the gathered tensor) and raises a per-rank
 * flag; qs_gather_wait enqueues a kernel that returns once all ranks have completed gather step `step_index` (1-based,
 * qs_gather_steps() = steps issued so far).  Tensor of step t: qs_gather_buffer(h, (t - 1) & 1); it stays valid until step t + 2. */
int qs_gather_create(QsHandle* h, int world_size, int rank, void* ipc_handle_out_64_bytes);
int qs_gather_connect(QsHandle* h, const void* ipc_handles_world_x_64_bytes);
void* qs_gather_buffer(QsHandle* h, int parity);
uint64_t qs_gather_steps(QsHandle* h);
int qs_gather_row_stride(QsHandle* h); /* floats between rows of the gathered tensors (D rounded up to 32: 128-byte aligned rows) */
int qs_gather_wait(QsHandle* h, uint64_t step_index, void* cuda_stream);
int qs_gather_close(QsHandle* h);

/* number of kernels this library has launched since create (bench.py "gpu_launches") */
int64_t qs_launch_count(QsHandle* h);

/* Diagnostics (scripts/ring_probe.py): device addresses of the finish-order queue ring ([8][num_envs] int32, see QsConfig.pipeline)
 * and of its publish counters ([8] uint32), the ring depth in use and the number of step launches issued so far.
 * Environment knobs read by the library, all for tests / experiments, none needed in production:
 *   QSTEP_GENERIC=1        run the generic (run-time dispatch) step kernel instead of the specialised variant
 *   QSTEP_RING_DEPTH=2..8  use fewer entries of the queue ring (read by qs_create)
 *   QSTEP_SEQ_START=<n>    start the launch sequence number at n (read by qs_create; reaches the 32-bit counter wrap quickly)
 *   QSTEP_QMAP=0|1         slot placement: 0 balanced over the CTAs, 1 finish-order groups
 *   QS_WARPS_PER_CTA=<n>   cap the warps per CTA (occupancy experiments) */
int qs_debug_queue(QsHandle* h, void** dev_queue, void** dev_publish_counters, int* ring_depth, uint64_t* step_launches);

#ifdef __cplusplus
}
#endif
#endif /* QSTEP_H_ */
