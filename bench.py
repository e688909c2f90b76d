#!/usr/bin/env python
"""Benchmark of the batched quadruped step (BASELINE.json metric: env-steps/s, mini_cheetah / flat / ALL_OBS,
4096 envs per GPU, random actions x50, auto-reset on termination).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the fp64 oracle port on all host cores

A "step" is one pass of the hot path over the whole batch: ONE launch of the fused step kernel, which also resets (in the
same warp) the envs that just terminated.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definition of every key.
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
    return {
        'workload': f'{ROBOT}/{SCENE}, ALL_OBS (D={OBS_DIM}), {envs} envs per GPU, ctrl = 50*N(0,1) per actuator, '
                    f'random-reset initial states, auto-reset on termination',
        'robot': ROBOT, 'scene': SCENE, 'envs_per_gpu': envs, 'global_envs': envs * n_gpus, 'obs_dim': OBS_DIM,
        'sim_dt': 0.002, 'algorithmic_bytes_per_env_step': BYTES_PER_ENV_STEP, 'parallelism': f'env-sharded x{n_gpus}',
    }


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    seed, n_envs, steps, duration = args
    import numpy as np

    from gym_quadruped_b200.model import Model
    from oracle.oracle import Oracle

    model = Model(ROBOT, SCENE)
    rng = np.random.RandomState(seed)
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
    for i in range(n_envs):
        e = Oracle(model)
        e.set_state(table[i % 32, :19], table[i % 32, 19:], np.zeros(18))
        e.set_env(rng.uniform(0.2, 1.5), -1.0, [rng.uniform(0.5, 1.0), 0, 0, 0])
        envs.append(e)
    chunk = 64
    ctrl = rng.randn(chunk, 12) * TORQUE_SCALE
    done_steps, cursor = 0, 0
    t0 = time.perf_counter()
    if steps is not None:  # fixed number of steps for every env
        for e in envs:
            left = steps
            while left > 0:
                k = min(chunk, left)
                _, cursor = e.rollout_autoreset(ctrl[:k], table, cursor)
                left -= k
            done_steps += steps
    else:  # run for a fixed duration
        while time.perf_counter() - t0 < duration:
            for e in envs:
                _, cursor = e.rollout_autoreset(ctrl, table, cursor)
                done_steps += chunk
    return done_steps, time.perf_counter() - t0


def cpu_throughput(steps_per_env=None, n_envs_total=None, duration=3.0, warmup=0):
    """Oracle env-steps/s on all host cores: one process per core, each advancing its share of the envs independently."""
    import multiprocessing as mp

    from oracle.oracle import build

    build()
    cores = os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        if steps_per_env is None:
            jobs = [(100 + i, 4, None, duration) for i in range(cores)]
        else:
            share = [n_envs_total // cores + (1 if i < n_envs_total % cores else 0) for i in range(cores)]
            if warmup:
                pool.map(_cpu_worker, [(900 + i, max(1, s), warmup, None) for i, s in enumerate(share)])
            jobs = [(100 + i, s, steps_per_env, None) for i, s in enumerate(share) if s > 0]
        res = pool.map(_cpu_worker, jobs)
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return total / wall, cores, total, wall


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU arm
    sample_envs = 512
    value, cores, total, wall = cpu_throughput(steps_per_env=args.steps, n_envs_total=sample_envs, warmup=args.warmup)
    sample = (f'{sample_envs} of {ENVS_PER_GPU} envs advanced {args.steps} steps each (after {args.warmup} warm-up steps) by the '
              f'fp64 C oracle port, one process per host core, auto-reset on termination; {total} env-steps in {wall:.2f} s')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(1, args.steps), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args.gpus, ENVS_PER_GPU),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'MuJoCo (the engine the reference calls) is not installable in this image; this is the from-scratch oracle port of its step',
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
    from gym_quadruped_b200.model import Model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; the CPU arm is `--impl reference`')
    torch.cuda.set_device(local_rank)
    dev = torch.device(f'cuda:{local_rank}')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world
    envs = args.envs
    K, W = args.steps, max(3, args.warmup)

    model = Model(ROBOT, SCENE)
    sim = BatchSim(model, envs, device=dev, seed=args.seed, env_id_offset=rank * envs, use_imu=USE_IMU, heightmap=HEIGHTMAP)
    opt = sim.make_reset_options(**RESET_KW)
    sim.reset(options=opt)

    # action ring larger than L2 (126 MB): 768 x 4096 x 12 fp32 = 151 MB per GPU, so ctrl is streamed from HBM every step
    ring = max(64, int(160e6 // (envs * 12 * 4)) + 1) if not args.small_ring else 64
    gen = torch.Generator(device=dev).manual_seed(args.seed + rank)
    actions = torch.randn(ring, envs, 12, device=dev, generator=gen) * TORQUE_SCALE

    def one_step(i):
        sim.step_autoreset(actions[i % ring], opt)  # one launch: step + in-kernel reset of the envs that just terminated

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        one_step(i)
    barrier()

    # ---- timed region: exactly K steps, device-timed, max over ranks
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        one_step(W + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    terminated_frac = float(sim.terminated.float().mean().item())

    # ---- dominant kernel alone: per-launch CUDA events on the launching stream (roofline.achieved)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        a, b = evs[i]
        a.record()
        one_step(W + K + i)
        b.record()
    torch.cuda.synchronize(dev)
    step_kernel_ms = sum(a.elapsed_time(b) for a, b in evs) / K

    # ---- same loop with the L2 flushed between iterations (256 MB write), each step timed on its own
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    kf = min(K, 200)
    cold = 0.0
    for i in range(kf):
        flush.fill_(i & 0xff)
        a, b = evs[i]
        a.record()
        one_step(i)
        b.record()
        torch.cuda.synchronize(dev)
        cold += a.elapsed_time(b)
    cold_ms = cold / kf
    del flush

    # ---- end to end through the C-ABI with HOST buffers (pinned): H2D ctrl + kernels + D2H obs/reward/flags every step
    host_ring = 8
    ctrl_h = [(torch.randn(envs, 12) * TORQUE_SCALE).pin_memory() for _ in range(host_ring)]
    obs_h = torch.empty(envs, sim.obs_dim).pin_memory(); rew_h = torch.empty(envs).pin_memory()
    term_h = torch.empty(envs, dtype=torch.uint8).pin_memory(); trunc_h = torch.empty(envs, dtype=torch.uint8).pin_memory()
    ke = min(K, 500)
    for i in range(3):
        sim.step_host(ctrl_h[i % host_ring], obs_h, rew_h, term_h, trunc_h, auto_reset=opt)
    barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        sim.step_host(ctrl_h[i % host_ring], obs_h, rew_h, term_h, trunc_h, auto_reset=opt)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    h2d = envs * 12 * 4
    d2h = envs * (sim.obs_dim * 4 + 4 + 1 + 1)

    # ---- optional: per-step NCCL all-gather of the observation tensor over NVLink (north_star's only collective)
    gather_ms = None
    if world > 1:
        gathered = torch.empty(world * envs, sim.obs_dim, device=dev)
        for i in range(3):
            one_step(i); dist.all_gather_into_tensor(gathered, sim.obs)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(K):
            one_step(i); dist.all_gather_into_tensor(gathered, sim.obs)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    # ---- reduce over ranks (max time)
    t = torch.tensor([ms, step_kernel_ms, cold_ms, e2e_s, gather_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, step_kernel_ms, cold_ms, e2e_s, gather_ms_max = [float(x) for x in t.tolist()]
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    if rank == 0:
        value = n_gpus * envs * K / (ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        achieved = envs * BYTES_PER_ENV_STEP / (step_kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / 'profiles' / 'traffic.json'
        if tp.exists() and args.workload == 'cfg2':
            try:
                traffic = json.loads(tp.read_text()).get('step_kernel_dram_bytes_per_launch')
            except Exception:
                traffic = None
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            v, cores, total, wall = cpu_throughput(duration=3.0)
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'fp64 C oracle port of the same workload, 4 envs per process x {cores} processes for {wall:.1f} s ({total} env-steps)'}
        cfg = workload_config(n_gpus, envs)
        cfg['l2'] = (f'inputs larger than L2: ctrl streamed from a {ring}-entry action ring ({ring * envs * 48 / 1e6:.0f} MB/GPU); env state '
                     f'({envs * 220 / 1e6:.1f} MB) is carried in place from step to step. L2-flushed variant reported as ms_per_step_l2_flushed')
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': peak_src, 'kernel': 'env_kernel<float,16,%d,MODE_STEP>' % (6 if ROBOT == 'go2' else 3), 'kernel_ms_per_launch': step_kernel_ms,
                         'note': 'latency/issue-bound path (about 1.4 kB and 60 kFLOP of dependent fp32 work per env-step): a low HBM fraction is expected'},
            'cpu_baseline': cpu,
            'e2e': {'value': n_gpus * envs * ke / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': ke,
                    'api': 'qs_step_host (C-ABI, pinned host buffers, in-kernel auto-reset)'},
            'gpu_launches': int(lt.item()), 'clocks': clocks,
            'ms_per_step_l2_flushed': cold_ms, 'terminated_fraction_last_step': terminated_frac,
        }
        if world > 1:
            line['with_obs_all_gather'] = {'value': n_gpus * envs * K / (gather_ms_max * 1e-3), 'unit': UNIT,
                                           'bytes_gathered_per_rank_per_step': world * envs * sim.obs_dim * 4}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=None, help='envs per GPU (default: the workload\'s)')
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS), help='cfg2 = BASELINE configs[1] (the bench line)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--small-ring', action='store_true', help='64-entry action ring (profiling runs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
