#!/usr/bin/env python
"""Benchmark of the batched quadruped step (BASELINE.json metric: env-steps/s, mini_cheetah / flat / ALL_OBS,
4096 envs per GPU, random actions x50, auto-reset on termination).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the fp64 oracle port on all host cores

A "step" is one pass of the hot path over the whole batch: ONE launch of the fused step kernel through the single-step C-ABI
call (qs_step_autoreset), which also resets (in the same warp) the envs that just terminated.  The timed region is exactly K such
launches between a barrier + synchronize on both sides; it is repeated (at least 5 times, and until 0.25 s have been timed) and the
MEDIAN region is reported, max over ranks.  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

ROBOT, SCENE = 'mini_cheetah', 'flat'
ENVS_PER_GPU = 4096
OBS_DIM = 227
BYTES_PER_ENV_STEP = 4 * (19 + 18 + 18 + 12) + 4 * (19 + 18 + 18) + 4 * OBS_DIM + 6  # 1402, SURVEY.md section 8(d)
TORQUE_SCALE = 50.0
RESET_KW = dict(lin_vel_range=(0.5, 1.0), ang_vel_range=(0.0, 0.0), friction_range=(0.2, 1.5), command_mode=1 | 4)
METRIC = 'env-steps/sec (total batch) mini_cheetah/flat @4096 envs/GPU'
UNIT = 'env-steps/s'
USE_IMU, HEIGHTMAP = False, None
PREROLL = 300  # untimed rollout steps before the warm-up: the timed steps see the steady-state mix of flight / contact / resets
# BASELINE.json configs[1] is the bench workload; configs[2..4] are parity-test cases (tests/test_gpu_parity.py) that can also
# be timed for information with --workload (per-GPU batch sizes of section 8(d))
WORKLOADS = {
    'cfg2': dict(robot='mini_cheetah', scene='flat', envs=4096, imu=False, hm=None),
    'cfg3': dict(robot='aliengo', scene='perlin', envs=8192, imu=False, hm=(5, 5, 0.1, 0.1)),
    'cfg4': dict(robot='go2', scene='random_boxes', envs=4096, imu=False, hm=None),
    'cfg5': dict(robot='hyqreal1', scene='flat', envs=8192, imu=True, hm=None),
}


def select_workload(name):
    global ROBOT, SCENE, ENVS_PER_GPU, OBS_DIM, BYTES_PER_ENV_STEP, METRIC, USE_IMU, HEIGHTMAP
    wl = WORKLOADS[name]
    ROBOT, SCENE, ENVS_PER_GPU, USE_IMU, HEIGHTMAP = wl['robot'], wl['scene'], wl['envs'], wl['imu'], wl['hm']
    OBS_DIM = 227 + (18 if USE_IMU else 0) + (HEIGHTMAP[0] * HEIGHTMAP[1] * 3 if HEIGHTMAP else 0)
    # section 8(d): state read + written, observation row, reward / flags (+ IMU bias read and written)
    BYTES_PER_ENV_STEP = 4 * (19 + 18 + 18 + 12) + 4 * (19 + 18 + 18) + 4 * OBS_DIM + 6 + (48 if USE_IMU else 0)
    if name != 'cfg2':
        METRIC = f'env-steps/sec (total batch) {ROBOT}/{SCENE} @{ENVS_PER_GPU} envs/GPU [informational workload {name}]'


def workload_config(n_gpus, envs):
    """Identical in both arms (the driver compares the two `config` objects)."""
    ring = max(64, int(160e6 // (envs * 12 * 4)) + 1)
    return {
        'workload': f'{ROBOT}/{SCENE}, ALL_OBS (D={OBS_DIM}), {envs} envs per GPU, ctrl = 50*N(0,1) per actuator, random-reset initial states '
                    f'advanced {PREROLL} untimed steps to the steady-state rollout, auto-reset on termination',
        'robot': ROBOT, 'scene': SCENE, 'envs_per_gpu': envs, 'global_envs': envs * n_gpus, 'obs_dim': OBS_DIM,
        'sim_dt': 0.002, 'algorithmic_bytes_per_env_step': BYTES_PER_ENV_STEP, 'parallelism': f'env-sharded x{n_gpus}',
        'l2': f'inputs larger than L2: ctrl is streamed from a {ring}-entry action ring ({ring * envs * 48 / 1e6:.0f} MB per GPU, L2 = 126 MB); '
              f'the env state ({envs * 220 / 1e6:.1f} MB) is carried in place from step to step, as in any rollout. An L2-flushed variant is '
              f'reported as ms_per_step_l2_flushed',
    }


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
_W = {}


def _cpu_init(seed0, n_envs_total, cores):
    """Pool initializer: every worker process builds its share of the envs once (not timed)."""
    import multiprocessing as mp

    import numpy as np

    from gym_quadruped_b200.model import Model
    from oracle.oracle import Oracle

    ident = mp.current_process()._identity
    wid = (ident[0] - 1) % cores if ident else 0
    share = n_envs_total // cores + (1 if wid < n_envs_total % cores else 0)
    model = Model(ROBOT, SCENE)
    rng = np.random.RandomState(seed0 + wid)
    key = np.array(model.c.key_qpos)
    # a table of lifted random-reset start states (reset distribution of quadruped_env.py:346-373, near the origin)
    o = Oracle(model)
    table = []
    for _ in range(32):
        q = key.copy()
        q[7:] += rng.uniform(-0.349, 0.349, 12)
        q[2] = model.hip_height
        r, p, y = rng.uniform(-0.1745, 0.1745), rng.uniform(-0.1745, 0.1745), rng.uniform(-np.pi, np.pi)
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
        q[3:7] = [cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy]
        o.set_state(q, np.zeros(18), np.zeros(18))
        o.lift()
        v = np.zeros(18); v[6:] = rng.uniform(-0.5, 0.5, 12)
        table.append(np.concatenate([o.get_state()[0], v]))
    table = np.array(table)
    envs = []
    for i in range(share):
        e = Oracle(model)
        e.set_state(table[i % 32, :19], table[i % 32, 19:], np.zeros(18))
        e.set_env(rng.uniform(0.2, 1.5), -1.0, [rng.uniform(0.5, 1.0), 0, 0, 0])
        envs.append(e)
    _W.update(envs=envs, table=table, cursor=0, rng=rng)


def _cpu_block(k):
    """Advance every env of this worker by k steps (random actions x50, auto-reset from the table); returns (env-steps, seconds)."""
    rng, table = _W['rng'], _W['table']
    ctrl = rng.randn(k, 12) * TORQUE_SCALE
    t0 = time.perf_counter()
    cur = _W['cursor']
    for e in _W['envs']:
        _, cur = e.rollout_autoreset(ctrl, table, cur)
    _W['cursor'] = cur
    return len(_W['envs']) * k, time.perf_counter() - t0


def cpu_throughput(steps, warmup, n_envs_total, min_seconds=1.0, preroll=PREROLL):
    """Oracle env-steps/s on all host cores: one process per core, each owning a fixed share of the `n_envs_total` envs.
    One block = every env advanced `steps` steps; blocks are repeated until `min_seconds` have been timed; the median block counts."""
    import multiprocessing as mp

    from oracle.oracle import build

    build()
    cores = os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(100, n_envs_total, cores)) as pool:
        def block(k):
            res = pool.map(_cpu_block, [k] * cores, chunksize=1)
            return sum(r[0] for r in res), max(r[1] for r in res)
        if preroll:
            block(preroll)
        if warmup:
            block(warmup)
        walls, total_steps, total_wall = [], 0, 0.0
        while len(walls) < 5 or total_wall < min_seconds:
            n, w = block(steps)
            walls.append(w); total_steps += n; total_wall += w
            if len(walls) >= 400:
                break
    wall = statistics.median(walls)
    return n_envs_total * steps / wall, cores, len(walls), wall, total_steps, total_wall


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU arm
    envs = args.envs
    value, cores, blocks, wall, total, total_wall = cpu_throughput(args.steps, args.warmup, envs, min_seconds=1.0)
    sample = (f'all {envs} envs of one GPU shard advanced by the fp64 C oracle port, one process per host core ({cores}), auto-reset on '
              f'termination; {PREROLL} untimed pre-roll + {args.warmup} warm-up steps, then {blocks} blocks of {args.steps} steps '
              f'({total} env-steps in {total_wall:.2f} s), median block')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(1, args.steps), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args.gpus, envs),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'MuJoCo (the engine the reference calls) is not installable in this image; this is the from-scratch oracle port of its step. '
                'The value is one shard of envs on the whole host; it does not scale with --gpus',
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        if not out.strip():  # the timed region was shorter than one sampling period: one query right after it
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-i', str(self.gpu)],
                                     capture_output=True, text=True, timeout=10).stdout
            except (OSError, subprocess.TimeoutExpired):
                out = ''
        for line in out.splitlines():
            p = [x.strip() for x in line.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_hbm_peak():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        try:
            return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from gym_quadruped_b200.backend import BatchSim
    from gym_quadruped_b200.distributed import ObsGather
    from gym_quadruped_b200.model import Model

    import faulthandler
    faulthandler.dump_traceback_later(420, exit=True)  # watchdog: a stuck collective must not hold the GPUs (stack traces, then exit)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; the CPU arm is `--impl reference`')
    torch.cuda.set_device(local_rank)
    dev = torch.device(f'cuda:{local_rank}')
    numa = ''
    if world > 1 and not os.environ.get('QS_BENCH_NO_NUMA'):
        from gym_quadruped_b200.distributed import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)  # pinned host buffers of the e2e path become local to this GPU's socket
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world
    envs = args.envs
    K, W = args.steps, max(3, args.warmup)

    model = Model(ROBOT, SCENE)
    sim_kw = dict(device=dev, seed=args.seed, env_id_offset=rank * envs, use_imu=USE_IMU, heightmap=HEIGHTMAP)
    # action ring larger than L2 (126 MB): 814 x 4096 x 12 fp32 = 160 MB per GPU, so ctrl is streamed from HBM every step
    ring = max(64, int(160e6 // (envs * 12 * 4)) + 1) if not args.small_ring else 64
    gen = torch.Generator(device=dev).manual_seed(args.seed + rank)
    actions = torch.randn(ring, envs, 12, device=dev, generator=gen) * TORQUE_SCALE

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_sim(pipeline, precision=0):
        s = BatchSim(model, envs, pipeline=pipeline, precision=precision, **sim_kw)
        o = s.make_reset_options(**RESET_KW)
        s.reset(options=o)
        for i in range(PREROLL):
            s.step_autoreset(actions[i % ring], o)
        return s, o

    def timed_regions(s, o, start, min_regions=5, min_seconds=0.25, max_regions=400):
        """Regions of exactly K single-step launches, each between barrier + synchronize; returns per-region ms (max over ranks)."""
        out, i, total = [], start, 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        while True:
            barrier()
            e0.record()
            for _ in range(K):
                s.step_autoreset(actions[i % ring], o)  # one launch: step + in-kernel reset of the envs that just terminated
                i += 1
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            stop = len(out) + 1 >= min_regions and total + ms >= 1e3 * min_seconds
            if world > 1:
                t = torch.tensor([ms, 0.0 if stop else 1.0], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, stop = float(t[0]), float(t[1]) == 0.0
            out.append(ms); total += ms
            if stop or len(out) >= max_regions:
                return out, i

    # ---- headline: pipelined single-step launches (QsConfig.pipeline: consecutive launches overlap on the device)
    sim, opt = make_sim(pipeline=not args.no_pipeline)
    for i in range(W):
        sim.step_autoreset(actions[(PREROLL + i) % ring], opt)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sim.launch_count
    regions, cursor = timed_regions(sim, opt, PREROLL + W)
    launches = (sim.launch_count - launches0) // len(regions)
    clocks = sampler.stop() if sampler else None
    ms = statistics.median(regions)
    torch.cuda.synchronize(dev)
    ncon = sim.ncon.cpu(); its = (sim.solver_iter.cpu() & 0xff)
    stats = {
        'mean_contacts_per_env': round(float(ncon.float().mean()), 3), 'envs_without_contact': round(float((ncon == 0).float().mean()), 3),
        'newton_iterations_mean': round(float(its.float().mean()), 3), 'newton_iterations_max': int(its.max()),
        'newton_iterations_hist': {str(k): round(float((its == k).float().mean()), 4) for k in range(0, 9)},
        'terminated_fraction_last_step': round(float(sim.terminated.float().mean()), 5), 'status_or': int(sim.status.max()),
    }
    variant = sim.step_variant

    # ---- the same launches in plain stream order (no overlap between consecutive launches), each launch timed on its own
    ser, ser_opt = make_sim(pipeline=False)
    for i in range(W):
        ser.step_autoreset(actions[(PREROLL + i) % ring], ser_opt)
    ser_regions, _ = timed_regions(ser, ser_opt, PREROLL + W, min_regions=5, min_seconds=0.1)
    ser_ms = statistics.median(ser_regions)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(K, 50))]
    for i, (a, b) in enumerate(evs):
        a.record()
        ser.step_autoreset(actions[i % ring], ser_opt)
        b.record()
    torch.cuda.synchronize(dev)
    launch_ms = statistics.median(a.elapsed_time(b) for a, b in evs)
    # qs_step_k: the same K launches issued by ONE library call on the non-pipelined handle (overlap without any caller contract)
    kreg = []
    ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(7):
        barrier()
        ek0.record()
        done = 0
        while done < K:  # exactly K steps; the action ring may be shorter than K
            base = (r * K + done) % ring
            cnt = min(K - done, ring - base)
            ser.step_k(actions[base:base + cnt], ser_opt)
            done += cnt
        ek1.record()
        barrier()
        kreg.append(ek0.elapsed_time(ek1))
    kstep_ms = statistics.median(kreg[2:])
    # L2 flushed between iterations (256 MB write), each step timed on its own
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    cold = []
    for i in range(min(len(evs), 100)):
        flush.fill_(i & 0xff)
        a, b = evs[i]
        a.record()
        ser.step_autoreset(actions[i % ring], ser_opt)
        b.record()
        torch.cuda.synchronize(dev)
        cold.append(a.elapsed_time(b))
    cold_ms = statistics.median(cold)
    del flush

    # ---- end to end through the C-ABI with HOST buffers (pinned): H2D ctrl + kernel + D2H obs/reward/flags every step, synchronous
    host_ring = 8
    ctrl_h = [(torch.randn(envs, 12) * TORQUE_SCALE).pin_memory() for _ in range(host_ring)]
    # rows padded to a multiple of 32 floats (128 B): every zero-copy row write crosses PCIe as whole lines (qs_step_host_strided)
    row = (sim.obs_dim + 31) // 32 * 32
    obs_h = torch.empty(envs, row).pin_memory()[:, :sim.obs_dim]; rew_h = torch.empty(envs).pin_memory()
    term_h = torch.empty(envs, dtype=torch.uint8).pin_memory(); trunc_h = torch.empty(envs, dtype=torch.uint8).pin_memory()
    for i in range(5):
        ser.step_host(ctrl_h[i % host_ring], obs_h, rew_h, term_h, trunc_h, auto_reset=ser_opt)
    e2e_regions, total = [], 0.0
    while True:
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            ser.step_host(ctrl_h[i % host_ring], obs_h, rew_h, term_h, trunc_h, auto_reset=ser_opt)
        dt = time.perf_counter() - t0
        e2e_regions.append(dt); total += dt
        more = (len(e2e_regions) < 5 or total < 0.25) and len(e2e_regions) < 400
        if world > 1:  # every rank must take the same number of trips through the barrier
            flag = torch.tensor([1.0 if more else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            more = bool(flag.item() > 0)
        if not more:
            break
    e2e_s = statistics.median(e2e_regions)
    h2d = envs * 12 * 4
    d2h = envs * (sim.obs_dim * 4 + 4 + 1 + 1)

    # ---- fp64 arithmetic build of the same kernel (QsConfig.precision = 1), for reference next to the fp32 value
    f64_ms = None
    if n_gpus == 1 and not args.no_fp64:
        s64, o64 = make_sim(pipeline=not args.no_pipeline, precision=1)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            s64.step_autoreset(actions[i % ring], o64)
        b.record()
        torch.cuda.synchronize(dev)
        f64_ms = a.elapsed_time(b) / K
        s64.close()

    # ---- the one collective north_star names: every rank ends up with the observation rows of all ranks
    gather = None
    if world > 1:
        gather = {}
        check = {}
        for mode in ('nccl', 'p2p'):
            gsim, gopt = make_sim(pipeline=not args.no_pipeline)  # same seed: both modes replay the same rollout
            try:
                g = ObsGather(gsim, mode=mode)
            except Exception as e:  # noqa: BLE001
                gather[mode] = {'unavailable': str(e)[:200]}
                continue

            def gstep(i):
                g.step_autoreset(actions[i % ring], gopt)
                with torch.cuda.stream(g.side):  # the consumer of the gathered rows lives on its own stream
                    rows = g.wait()  # (an empty consumer: the double buffer is free again long before step t + 2)
                return rows
            for i in range(10):
                rows = gstep(i)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reg = []
            for _ in range(5):
                barrier()
                g0.record()
                for i in range(K):
                    rows = gstep(10 + i)
                torch.cuda.current_stream(dev).wait_stream(g.side)  # the region ends when the last gathered tensor is complete
                g1.record()
                barrier()
                reg.append(g0.elapsed_time(g1))
            check[mode] = rows.clone()
            t = torch.tensor(reg, dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gather[mode] = {'ms_per_step': float(t.median()) / K, 'value': n_gpus * envs * K / (float(t.median()) * 1e-3),
                            'bytes_received_per_rank_per_step': (world - 1) * envs * gsim.obs_dim * 4, 'how': g.describe()}
            g.close()
            gsim.close()
        if len(check) == 2:  # both modes ran the same rollout: the gathered tensors must be bit-identical, on every rank
            same = torch.tensor([float(torch.equal(check['nccl'], check['p2p']))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            gather['p2p_equals_nccl_bitwise'] = bool(same.item() == 1.0)

    # ---- reduce over ranks (max time)
    t = torch.tensor([ser_ms, launch_ms, cold_ms, e2e_s, f64_ms or 0.0, kstep_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ser_ms, launch_ms, cold_ms, e2e_s, f64_ms_r, kstep_ms = [float(x) for x in t.tolist()]
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    if rank == 0:
        value = n_gpus * envs * K / (ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        achieved = envs * BYTES_PER_ENV_STEP / (ms / K * 1e-3) / 1e9
        traffic = None
        tp = ROOT / 'profiles' / 'traffic.json'
        if tp.exists() and args.workload == 'cfg2':
            try:
                traffic = json.loads(tp.read_text()).get('step_kernel_dram_bytes_per_launch')
            except Exception:
                traffic = None
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            v, cores, blocks, wall, total, total_wall = cpu_throughput(20, 5, envs, min_seconds=1.2)
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'fp64 C oracle port, all {envs} envs on {cores} processes (one per host core), {blocks} blocks of 20 steps after '
                             f'{PREROLL}+5 untimed steps ({total} env-steps in {total_wall:.1f} s), median block'}
        e2e_ms = 1e3 * e2e_s / K
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(n_gpus, envs),
            'timing': {'regions': len(regions), 'region_ms_median': ms, 'region_ms_min': min(regions), 'region_ms_max': max(regions),
                       'launch_mode': 'serialized' if args.no_pipeline else
                       'pipelined: QsConfig.pipeline=1, every step is its own qs_step_autoreset launch; consecutive launches overlap on the '
                       'device (programmatic dependent launch + per-env finish-order queues), results bit-identical to the serialized order',
                       'step_kernel_variant': variant},
            'serialized': {'value': n_gpus * envs * K / (ser_ms * 1e-3), 'ms_per_step': ser_ms / K, 'ms_per_launch_events': launch_ms,
                           'note': 'same launches without overlap (QsConfig.pipeline=0): what a caller gets whose next action depends on '
                                   'the previous observation'},
            'k_step_call': {'value': n_gpus * envs * K / (kstep_ms * 1e-3), 'ms_per_step': kstep_ms / K,
                            'note': 'qs_step_k: K steps from one C-ABI call on the pipeline=0 handle, one launch per step, overlapped by the library; '
                                    'per-step outputs bit-identical to K single calls (tests/test_gpu_round2.py::test_step_k_equals_k_single_steps)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': peak_src, 'kernel': f'env_kernel<float,16,{6 if ROBOT == "go2" else 3},MODE_STEP,{variant}>',
                         'kernel_ms_per_launch': ms / K, 'kernel_ms_per_launch_serialized': launch_ms,
                         'note': 'latency/issue-bound path (about 1.4 kB and 60 kFLOP of dependent fp32 work per env-step): a low HBM '
                                 'fraction is expected; achieved = algorithmic bytes per launch / (timed region / K launches)'},
            'cpu_baseline': cpu,
            'e2e': {'value': n_gpus * envs * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': K,
                    'regions': len(e2e_regions), 'ms_per_step': e2e_ms,
                    'api': 'qs_step_host_strided (C-ABI, pinned host buffers, observation rows on 128-byte boundaries, in-kernel auto-reset), '
                           'synchronous: returns when the results are in host memory',
                    'exposed_transfer_ms_per_step': e2e_ms - ser_ms / K, 'numa_binding_rank0': numa,
                    'note': 'zero-copy: the kernel reads ctrl and writes obs rows through the mapped host buffers; exposed_transfer = e2e - '
                            'serialized device time per step is what PCIe adds on top of the kernel'},
            'gpu_launches': int(lt.item()), 'clocks': clocks,
            'ms_per_step_l2_flushed': cold_ms, 'workload_stats': stats,
            'fp64_arithmetic': None if f64_ms is None else {'value': envs * K / (f64_ms_r * K * 1e-3), 'ms_per_step': f64_ms_r,
                                                            'note': 'QsConfig.precision=1: same kernel source in fp64 arithmetic (8 warps per CTA)'},
        }
        if gather is not None:
            line['with_obs_all_gather'] = gather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)   # >= 200 warm-up, >= 1000 timed steps, median of 5
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=None, help='envs per GPU (default: the workload\'s)')
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS), help='cfg2 = BASELINE configs[1] (the bench line)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--small-ring', action='store_true', help='64-entry action ring (profiling runs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fp64', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true', help='headline in plain stream order (QsConfig.pipeline=0)')
    args = ap.parse_args()
    select_workload(args.workload)
    if args.envs is None:
        args.envs = ENVS_PER_GPU
    if args.impl == 'reference':
        run_reference(args)
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one process per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', str(29500 + os.getpid() % 2000), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == '__main__':
    main()
