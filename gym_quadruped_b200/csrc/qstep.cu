// qstep.cu -- libqstep: the C-ABI of include/qstep.h (host side).  The kernels live in qs_kernel.cuh / qs_env.cuh and are
// instantiated, one variant per translation unit, in qs_inst_*.cu; this file owns the handle, picks the kernel variant for a
// model, and enqueues launches.
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
#include "qs_kernel.cuh"
#include "qs_variants.h"

using namespace qs;

namespace {

const VariantInfo* find_variant(int precision, int maxdim, int have, bool generic_only) {
  const VariantInfo* best = nullptr;
  int best_bits = -1;
  for (const VariantInfo& v : all_variants()) {
    if (v.precision != precision || v.maxdim != maxdim) continue;
    if (generic_only ? v.feat != 0 : (v.feat & ~have) != 0) continue;
    const int bits = __builtin_popcount(unsigned(v.feat));
    if (bits > best_bits) { best = &v; best_bits = bits; }
  }
  return best;
}

// initial draw of the disturbance schedule for every env (quadruped_env.py:240-242 calls _sample_external_disturbances in __init__)
__global__ void schedule_init_kernel(const KParams p) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.num_envs) return;
  const unsigned env_g = unsigned(env + p.env_id_offset);
  p.b.cmd_count[env] = 0; p.b.cmd_limit[env] = 0x7fffffff;  // the first reset draws the limit (:1068-1070)
  p.b.ext_count[env] = 0;
  if (!p.sch_ext_enabled) {
    p.b.ext_limit[env] = 0x7fffffff;
    for (int i = 0; i < 6; i++) p.b.ext_wrench[size_t(env) * 6 + i] = 0.f;
    return;
  }
  const unsigned ee = p.ext_epoch[env];
  for (int i = 0; i < 7; i++) {
    uint32_t r[4];
    philox4x32(env_g, ee, unsigned(i), 0xD157u, p.seed_lo ^ 0x7F4A7C15u, p.seed_hi, r);
    const float u = u32_to_unit(r[0]);
    if (i < 6) p.b.ext_wrench[size_t(env) * 6 + i] = float(double(p.sch_ext_lo[i]) + (double(p.sch_ext_hi[i]) - double(p.sch_ext_lo[i])) * double(u));
    else p.b.ext_limit[env] = 1000 + int(u * 2000.f);
  }
  p.ext_epoch[env] = ee + 1;
}

// qs_gather_wait: one thread per rank polls this rank's flag slots until every rank has published step `seq` of this parity
__global__ void gather_wait_kernel(const unsigned* flags, int world, unsigned seq) {
  if (int(threadIdx.x) < world) {
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
      if (int(v - seq) < 0) __nanosleep(200);
    } while (int(v - seq) < 0);
  }
}

__global__ void queue_init_kernel(int* q, unsigned* tails, int n, unsigned tail0, int gen0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    q[i] = (gen0 << QS_SLOT_ENV_BITS) | i;  // ring entry 0: identity placement for the first launch (generation gen0)
    for (int r = 1; r < QS_QUEUE_DEPTH; r++) q[size_t(r) * n + i] = -1 - ((gen0 - 1) & QS_SLOT_GEN_MASK);  // empty, first read is generation gen0
  }
  if (i < QS_QUEUE_DEPTH) tails[i] = tail0;
}

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
  int* d_queue = nullptr;          // QS_QUEUE_DEPTH x [N] finish-order queues (see KParams)
  unsigned* d_queue_tail = nullptr; // one monotonic publish counter per ring entry
  int ring_depth = QS_QUEUE_DEPTH;  // ring entries in use (QSTEP_RING_DEPTH lowers it: tests of the depth-limited case)
  uint64_t step_seq = 0;           // number of step launches so far: launch s reads ring entry s % DEPTH and fills (s + 1) % DEPTH
  bool last_was_step = false;      // the previous launch of this handle was a step kernel (nothing of ours in between)
  void* last_stream = nullptr;
  const char* step_variant = "";
  // peer-to-peer observation gather (qs_gather_*)
  struct {
    int world = 0, rank = 0;
    unsigned char* block = nullptr;     // this rank's allocation: 2 gathered tensors + 2 x 8 flags
    unsigned char* peer[8] = {};        // the same block on every rank (peer[rank] == block)
    size_t tensor_bytes = 0;
    int row_stride = 0;                 // floats per row of the gathered tensors: D rounded up to 32 (rows start on 128-byte boundaries)
    uint64_t steps = 0;                 // gather steps launched so far
    bool connected = false;
  } gather;
  QsSchedule sched{};
  unsigned* d_episode = nullptr;
  unsigned* d_tick = nullptr;
  unsigned* d_cmd_epoch = nullptr;
  unsigned* d_ext_epoch = nullptr;
  uint64_t seed = 0;
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

template <typename real> static int setup_variant(QsHandle* h, const QsModel* model) {
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
  const int precision = sizeof(real) == 4 ? 0 : 1;
  const VariantInfo* gen = find_variant(precision, h->maxdim, 0, true);
  if (!gen) return fail(h, 4, "no generic kernel variant compiled for this precision / contact dimension");
  const VariantInfo* gen3 = find_variant(precision, 3, 0, true);
  // QSTEP_GENERIC=1 forces the run-time-dispatch kernel (tests compare it bit for bit with the specialised variants)
  const char* force = getenv("QSTEP_GENERIC");
  const VariantInfo* spec = (force && force[0] == '1') ? gen : find_variant(precision, h->maxdim, model_features(*model, h->cfg.use_imu != 0, h->cfg.hm_rows * h->cfg.hm_cols), false);
  h->k_step = spec->step; h->k_reset = gen->reset; h->k_forward = gen->forward;
  h->k_raycast = gen3 ? gen3->raycast : nullptr;
  h->step_variant = spec->name;
  int dev = 0, max_smem = 0;
  QS_CUDA(h, cudaGetDevice(&dev));
  QS_CUDA(h, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int warps = gen->max_warps;
  if (const char* ev = getenv("QS_WARPS_PER_CTA")) { const int v = atoi(ev); if (v >= 1 && v < warps) warps = v; }  // occupancy experiments
  while (warps > 1 && gen->dm_bytes + 128 + warps * gen->ws_bytes > size_t(max_smem)) warps--;
  h->warps_per_cta = warps;
  h->smem_bytes = gen->dm_bytes + 128 + warps * gen->ws_bytes;
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

static KParams base_params(QsHandle* h);

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
// diagnostics (scripts/ring_probe.py): device addresses of the finish-order queue ring and its publish counters, ring depth, launch count
int qs_debug_queue(QsHandle* h, void** queue, void** tails, int* depth, uint64_t* step_seq) {
  if (!h) return 1;
  *queue = h->d_queue; *tails = h->d_queue_tail; *depth = h->ring_depth; *step_seq = h->step_seq;
  return 0;
}

int qs_create(const QsModel* model, const QsConfig* cfg, QsHandle** out) {
  if (!model || !cfg || !out) return fail(nullptr, 1, "null argument");
  if (cfg->num_envs <= 0) return fail(nullptr, 1, "num_envs must be positive");
  if (cfg->num_envs >= (1 << QS_SLOT_ENV_BITS)) return fail(nullptr, 1, "num_envs must be below 1048576 per handle");
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
  h->seed = cfg->seed;
  if (const char* ev = getenv("QSTEP_RING_DEPTH")) { const int v = atoi(ev); if (v >= 2 && v <= QS_QUEUE_DEPTH) h->ring_depth = v; }
  rc = cfg->precision == 0 ? setup_variant<float>(h, model) : setup_variant<double>(h, model);
  if (rc == 0) {
    const size_t n = size_t(cfg->num_envs);
    unsigned* ctr = nullptr;  // episode | tick | cmd_epoch | ext_epoch
    cudaError_t e1 = cudaMalloc(&ctr, 4 * n * sizeof(unsigned)), e2 = cudaMalloc(&h->d_queue, QS_QUEUE_DEPTH * n * sizeof(int)),
                e3 = cudaMalloc(&h->d_queue_tail, QS_QUEUE_DEPTH * sizeof(unsigned));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) rc = fail(h, 5, "cudaMalloc failed");
    else {
      h->d_episode = ctr; h->d_tick = ctr + n; h->d_cmd_epoch = ctr + 2 * n; h->d_ext_epoch = ctr + 3 * n;
      cudaMemset(ctr, 0, 4 * n * sizeof(unsigned));
      // QSTEP_SEQ_START: pretend that many step launches have already happened (rounded down to a multiple of the ring depth), so
      // that tests reach the 32-bit wrap of the publish counters (4096 envs: every 8.4 M launches) within a few steps
      if (const char* ev = getenv("QSTEP_SEQ_START")) h->step_seq = strtoull(ev, nullptr, 10) / uint64_t(h->ring_depth) * uint64_t(h->ring_depth);
      const uint64_t g0 = h->step_seq / uint64_t(h->ring_depth);
      queue_init_kernel<<<unsigned((n + 255) / 256), 256>>>(h->d_queue, h->d_queue_tail, int(n), unsigned(g0 * n), int(g0 & QS_SLOT_GEN_MASK));
      if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(h, 5, "queue initialisation failed");
    }
  }
  if (rc != 0) { g_create_error = h->err; qs_destroy(h); return rc; }
  *out = h;
  return 0;
}

void qs_destroy(QsHandle* h) {
  if (!h) return;
  if (h->gather.block) qs_gather_close(h);
  cudaFree(h->d_queue); cudaFree(h->d_queue_tail);
  cudaFree(h->d_dm); cudaFree(h->d_vert); cudaFree(h->d_hf); cudaFree(h->d_boxes); cudaFree(h->d_episode); cudaFree(h->d_aux);
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
  h->last_was_step = false;
  return qs_set_schedule(h, nullptr, nullptr);  // counters / limits of the new buffers start from a defined state
}

int qs_set_schedule(QsHandle* h, const QsSchedule* sched, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_set_schedule: handle not bound");
  h->sched = sched ? *sched : QsSchedule{};
  if (sched) {
    for (int i = 0; i < 6; i++)
      if (!(sched->ext_lo[i] == sched->ext_lo[i]) || !(sched->ext_hi[i] == sched->ext_hi[i])) return fail(h, 1, "qs_set_schedule: NaN range");
  }
  KParams p = base_params(h);
  h->last_was_step = false;
  schedule_init_kernel<<<(h->cfg.num_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

int qs_set_seed(QsHandle* h, uint64_t seed, void* stream) {
  if (!h) return fail(h, 1, "null argument");
  h->seed = seed;
  h->last_was_step = false;
  QS_CUDA(h, cudaMemsetAsync(h->d_episode, 0, 4 * size_t(h->cfg.num_envs) * sizeof(unsigned), static_cast<cudaStream_t>(stream)));
  return 0;
}

const char* qs_step_variant(QsHandle* h) { return h ? h->step_variant : ""; }

int qs_gather_create(QsHandle* h, int world, int rank, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail(h, 1, "qs_gather_create: null argument");
  if (world < 2 || world > 8 || rank < 0 || rank >= world) return fail(h, 1, "qs_gather_create: world must be 2..8");
  if (h->gather.block) return fail(h, 1, "qs_gather_create: already created");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  auto& g = h->gather;
  g.world = world; g.rank = rank;
  g.row_stride = (h->obs_dim + 31) / 32 * 32;
  g.tensor_bytes = (size_t(world) * h->cfg.num_envs * g.row_stride * sizeof(float) + 255) & ~size_t(255);
  const size_t total = 2 * g.tensor_bytes + 256;
  QS_CUDA(h, cudaMalloc(&g.block, total));
  QS_CUDA(h, cudaMemset(g.block, 0, total));
  cudaIpcMemHandle_t hd;
  QS_CUDA(h, cudaIpcGetMemHandle(&hd, g.block));
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  return 0;
}

int qs_gather_connect(QsHandle* h, const void* ipc_handles) {
  if (!h || !h->gather.block || !ipc_handles) return fail(h, 1, "qs_gather_connect: call qs_gather_create first");
  auto& g = h->gather;
  for (int q = 0; q < g.world; q++) {
    if (q == g.rank) { g.peer[q] = g.block; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const unsigned char*>(ipc_handles) + 64 * q, sizeof(hd));
    void* ptr = nullptr;
    QS_CUDA(h, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    g.peer[q] = static_cast<unsigned char*>(ptr);
  }
  g.connected = true;
  h->last_was_step = false;
  return 0;
}

void* qs_gather_buffer(QsHandle* h, int parity) {
  return (h && h->gather.block) ? h->gather.block + size_t(parity & 1) * h->gather.tensor_bytes : nullptr;
}

uint64_t qs_gather_steps(QsHandle* h) { return h ? h->gather.steps : 0; }
int qs_gather_row_stride(QsHandle* h) { return h ? h->gather.row_stride : 0; }

int qs_gather_wait(QsHandle* h, uint64_t step_index, void* stream) {
  if (!h || !h->gather.connected) return fail(h, 1, "qs_gather_wait: gather not connected");
  if (step_index == 0) return 0;
  auto& g = h->gather;
  const int k = int((step_index - 1) & 1);
  const unsigned seq = unsigned((step_index + 1) / 2);
  const unsigned* flags = reinterpret_cast<const unsigned*>(g.block + 2 * g.tensor_bytes) + 8 * k;
  gather_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, g.world, seq);
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

int qs_gather_close(QsHandle* h) {
  if (!h) return 1;
  auto& g = h->gather;
  for (int q = 0; q < g.world; q++)
    if (q != g.rank && g.peer[q]) cudaIpcCloseMemHandle(g.peer[q]);
  if (g.block) cudaFree(g.block);
  g = {};
  h->last_was_step = false;
  return 0;
}

static KParams base_params(QsHandle* h) {
  KParams p{};
  p.dm = h->d_dm; p.vert = h->d_vert; p.hf = h->d_hf; p.boxes = h->d_boxes;
  p.hm_rows = h->cfg.hm_rows; p.hm_cols = h->cfg.hm_cols; p.hm_dx = float(h->cfg.hm_dx); p.hm_dy = float(h->cfg.hm_dy);
  p.num_envs = h->cfg.num_envs; p.obs_dim = h->obs_dim; p.obs_stride = h->obs_dim; p.use_imu = h->cfg.use_imu;
  p.max_iter = h->cfg.solver_max_iter > 0 ? h->cfg.solver_max_iter : (h->cfg.precision == 0 ? 50 : 100);
  p.tol = h->cfg.precision == 0 ? 1e-6f : 1e-8f;
  p.env_id_offset = h->cfg.env_id_offset;
  p.seed_lo = unsigned(h->seed & 0xffffffffu); p.seed_hi = unsigned(h->seed >> 32);
  p.sch_command_mode = h->sched.command_mode; p.sch_ext_enabled = h->sched.ext_enabled;
  for (int i = 0; i < 2; i++) { p.sch_lin[i] = float(h->sched.lin_vel_range[i]); p.sch_ang[i] = float(h->sched.ang_vel_range[i]); }
  for (int i = 0; i < 6; i++) { p.sch_ext_lo[i] = float(h->sched.ext_lo[i]); p.sch_ext_hi[i] = float(h->sched.ext_hi[i]); }
  p.cmd_epoch = h->d_cmd_epoch; p.ext_epoch = h->d_ext_epoch;
  p.imu_an = float(h->cfg.imu_accel_noise); p.imu_gn = float(h->cfg.imu_gyro_noise);
  p.imu_abr = float(h->cfg.imu_accel_bias_rate); p.imu_gbr = float(h->cfg.imu_gyro_bias_rate);
  p.b = h->buf; p.episode = h->d_episode; p.tick = h->d_tick;
  return p;
}

// `chained`: this step launch directly follows another step launch of the same handle on the same stream and the handle allows
// overlap (QsConfig.pipeline): it is enqueued with programmatic stream serialization, so its CTAs may start while the previous
// launch is still running; the per-env dependency is carried by the finish-order queues, not by the grid boundary.
static int launch(QsHandle* h, KernelFn fn, const KParams& p, cudaStream_t s, bool chained = false) {
  const int warps = h->warps_per_cta, grid = (h->cfg.num_envs + warps - 1) / warps;
  if (chained) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(unsigned(grid)); lc.blockDim = dim3(unsigned(warps * 32)); lc.dynamicSmemBytes = h->smem_bytes; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    QS_CUDA(h, cudaLaunchKernelEx(&lc, fn, p));
  } else {
    fn<<<grid, warps * 32, h->smem_bytes, s>>>(p);
  }
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

// kmode 0: plain call (overlap with the previous launch only if QsConfig.pipeline allows it); 1 / 2: first / later launch of a
// qs_step_k sequence -- the library itself issues these launches back to back, so the later ones may always overlap their predecessor
static int step_impl(QsHandle* h, const float* ctrl, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated,
                     const QsResetOptions* auto_reset, void* stream, int kmode = 0, size_t obs_row_stride = 0) {
  if (!h || !h->bound) return fail(h, 1, "qs_step: handle not bound");
  if (!ctrl) return fail(h, 1, "qs_step: ctrl is null");
  KParams p = base_params(h);
  p.ctrl = ctrl; p.obs = obs; p.reward = reward; p.terminated = terminated; p.truncated = truncated;
  if (obs_row_stride) p.obs_stride = int(obs_row_stride);
  if (auto_reset) { p.auto_reset = 1; p.ro = *auto_reset; }
#ifdef QS_PROF
  p.prof = h->prof;
#endif
  {
    const size_t n = size_t(h->cfg.num_envs);
    const uint64_t s = h->step_seq++;
    const uint64_t D = uint64_t(h->ring_depth);
    const int in = int(s % D), out = int((s + 1) % D);
    p.q_in = h->d_queue + size_t(in) * n; p.q_out = h->d_queue + size_t(out) * n;
    p.q_tail = h->d_queue_tail + out;
    p.q_gen_in = int((s / D) & QS_SLOT_GEN_MASK); p.q_gen_out = int(((s + 1) / D) & QS_SLOT_GEN_MASK);
    p.q_tail_base = unsigned((s / D) * n);  // launches s - DEPTH, s - 2 DEPTH, ... filled this ring entry before: n envs each
    p.q_contiguous = (h->cfg.pipeline || kmode) ? 1 : 0;
    if (const char* ev = getenv("QSTEP_QMAP")) p.q_contiguous = atoi(ev);  // placement experiments: 0 balanced, 1 finish-order groups
    p.q_sync = (h->cfg.pipeline || kmode) ? 1 : 0;
  }
  if (h->gather.connected) {
    auto& g = h->gather;
    const int k = int(g.steps & 1);
    g.steps++;
    p.gather_world = g.world; p.gather_rank = g.rank;
    p.gather_seq = unsigned((g.steps + 1) / 2);  // 1, 1, 2, 2, ...: the count of steps of this parity
    for (int q = 0; q < g.world; q++) {
      p.gather_peers[q] = reinterpret_cast<float*>(g.peer[q] + size_t(k) * g.tensor_bytes);
      p.gather_flags[q] = reinterpret_cast<unsigned*>(g.peer[q] + 2 * g.tensor_bytes) + 8 * k;
    }
    p.obs = p.gather_peers[g.rank] + size_t(g.rank) * h->cfg.num_envs * g.row_stride;  // own rows live in the gathered tensor itself
    p.obs_stride = g.row_stride;
  }
  const bool chained = kmode == 2 || (kmode == 0 && h->cfg.pipeline && h->last_was_step && h->last_stream == stream);
  const int rc = launch(h, h->k_step, p, static_cast<cudaStream_t>(stream), chained);
  h->last_was_step = rc == 0; h->last_stream = stream;
  return rc;
}

int qs_step_k(QsHandle* h, int k, const float* ctrl, const QsResetOptions* auto_reset, float* obs, size_t obs_step_stride, float* reward,
              uint8_t* terminated, uint8_t* truncated, void* stream) {
  if (k <= 0) return fail(h, 1, "qs_step_k: k must be positive");
  if (!h) return fail(h, 1, "null handle");
  const size_t n = size_t(h->cfg.num_envs);
  for (int i = 0; i < k; i++) {
    const int rc = step_impl(h, ctrl + size_t(i) * n * NU, obs ? obs + size_t(i) * obs_step_stride : nullptr, reward ? reward + size_t(i) * n : nullptr,
                             terminated ? terminated + size_t(i) * n : nullptr, truncated ? truncated + size_t(i) * n : nullptr, auto_reset,
                             stream, i == 0 ? 1 : 2);
    if (rc) return rc;
  }
  if (!h->cfg.pipeline) h->last_was_step = false;  // a later plain qs_step of a non-pipelined handle keeps full stream order
  return 0;
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
  return qs_step_host_strided(h, ctrl, auto_reset, obs, 0, reward, terminated, truncated, stream);
}

int qs_step_host_strided(QsHandle* h, const float* ctrl, const QsResetOptions* auto_reset, float* obs, size_t obs_row_stride, float* reward,
                         uint8_t* terminated, uint8_t* truncated, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_step_host: handle not bound");
  if (obs_row_stride && obs_row_stride < size_t(h->obs_dim)) return fail(h, 1, "qs_step_host_strided: row stride smaller than the observation width");
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
  if (obs && !k_obs && obs_row_stride && obs_row_stride != size_t(h->obs_dim)) return fail(h, 1, "qs_step_host_strided: padded rows need a pinned (mapped) observation buffer");
  int rc = step_impl(h, k_ctrl ? k_ctrl : h->d_ctrl, obs ? (k_obs ? k_obs : h->d_obs) : nullptr, reward ? (k_rew ? k_rew : h->d_reward) : nullptr,
                     terminated ? (k_term ? k_term : h->d_term) : h->d_term, truncated ? (k_trunc ? k_trunc : h->d_trunc) : nullptr,
                     auto_reset, stream, 0, k_obs ? obs_row_stride : 0);
  if (rc) return rc;
  if (obs && !k_obs) QS_CUDA(h, cudaMemcpyAsync(obs, h->d_obs, n * h->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (reward && !k_rew) QS_CUDA(h, cudaMemcpyAsync(reward, h->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (terminated && !k_term) QS_CUDA(h, cudaMemcpyAsync(terminated, h->d_term, n, cudaMemcpyDeviceToHost, s));
  if (truncated && !k_trunc) QS_CUDA(h, cudaMemcpyAsync(truncated, h->d_trunc, n, cudaMemcpyDeviceToHost, s));
  QS_CUDA(h, cudaStreamSynchronize(s));
  h->last_was_step = false;  // the caller rewrites the host buffers between calls: every host-buffer step is a chain of one
  return 0;
}

static int reset_impl(QsHandle* h, const uint8_t* mask, const float* qpos, const float* qvel, const QsResetOptions* opt, float* obs, void* stream) {
  if (!h || !h->bound) return fail(h, 1, "qs_reset: handle not bound");
  if (!opt) return fail(h, 1, "qs_reset: options are null");
  if ((qpos == nullptr) != (qvel == nullptr)) return fail(h, 1, "qs_reset: qpos and qvel must be given together");
  KParams p = base_params(h);
  p.mask = mask; p.in_qpos = qpos; p.in_qvel = qvel; p.ro = *opt; p.obs = obs;
  h->last_was_step = false;
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
  h->last_was_step = false;
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
  if (!h->k_raycast) return fail(h, 4, "ray-cast kernel not compiled");
  h->last_was_step = false;
  h->k_raycast<<<grid, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  h->launches++;
  QS_CUDA(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
