"""Env sharding across GPUs: contiguous env-index ranges per rank, no data-path collective.  The only
collective north_star names is the gather of the observation rows, `ObsGather`:

* mode 'p2p'  -- fused into the step kernel: every warp stores its finished observation row straight into the gathered tensor of
  EVERY rank through peer-mapped memory (CUDA IPC over NVLink / NVSwitch), the last warp of the launch raises a per-rank flag on
  every peer; no separate collective kernel, the transfer overlaps the envs that are still being solved (csrc/qs_kernel.cuh).
* mode 'nccl' -- `all_gather_into_tensor` on a side stream, double-buffered so that it overlaps the next step (the baseline).

`gather_rows` is the plain blocking helper (gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def shard_range(rank: int, world_size: int, total_envs: int) -> tuple[int, int]:
    """[start, stop) of the global env ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally sized [n_local, D] row blocks into [world * n_local, D], rank-major (global env order)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class _DevArray:
    """Wraps a raw device pointer owned by libqstep so that torch can view it (`__cuda_array_interface__`)."""

    def __init__(self, ptr: int, shape, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


class ObsGather:
    """Steps a `BatchSim` shard and keeps, on every rank, the observation rows of ALL ranks: `gathered[k]` is a
    [world * N, D] tensor in global env order, k = step parity (double buffer: the rows of step t stay valid until step t + 2 is
    enqueued).  Per step:

        g.step_autoreset(ctrl, opt)              # on the current stream; in 'p2p' mode the rows travel from inside the kernel
        with torch.cuda.stream(g.side):          # the consumer lives on its own stream, so the next step is not held back
            rows = g.wait()                      # stream-ordered: work enqueued after it sees the complete tensor of the last step
            ... consume rows ...

    Double-buffer contract: the tensor of step t is overwritten by step t + 2, on every rank.  A closed-loop caller (the actions of
    step t + 1 are computed from `rows`) satisfies it by construction.  A consumer that may lag more than one step behind calls
    `g.release()` after consuming: step t + 2 then waits for it -- at the price of an event wait between two step launches,
    which ends their overlap on the device (QsConfig.pipeline)."""

    def __init__(self, sim, mode: str = 'p2p', group=None):
        if not dist.is_initialized():
            raise RuntimeError('ObsGather needs an initialised torch.distributed process group')
        self.sim, self.mode, self.group = sim, mode, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.N, self.D, self.dev = sim.N, sim.obs_dim, sim.device
        self.t = 0
        self.side = torch.cuda.Stream(device=self.dev)
        self._coll = torch.cuda.Stream(device=self.dev)  # 'nccl' mode: the collective's stream
        self._ev_step = [torch.cuda.Event() for _ in range(2)]
        self._ev_gath = [torch.cuda.Event() for _ in range(2)]
        self._ev_rel = [torch.cuda.Event() for _ in range(2)]
        self._rel_pending = [False, False]
        if mode == 'nccl':
            self.local = [torch.zeros(self.N, self.D, device=self.dev) for _ in range(2)]
            self.gathered = [torch.zeros(self.world * self.N, self.D, device=self.dev) for _ in range(2)]
        elif mode == 'p2p':
            self._open_p2p()
        else:
            raise ValueError(f'unknown gather mode {mode}')

    # ------------------------------------------------------------------ p2p plumbing (CUDA IPC handles exchanged through the process group)
    def _open_p2p(self):
        L, sim = self.sim.L, self.sim
        if self.world > 8:
            raise RuntimeError('the peer-to-peer gather spans the GPUs of one node (at most 8)')
        L.qs_gather_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.qs_gather_connect.argtypes = [C.c_void_p, C.c_void_p]
        L.qs_gather_buffer.argtypes = [C.c_void_p, C.c_int]
        L.qs_gather_buffer.restype = C.c_void_p
        L.qs_gather_wait.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.qs_gather_steps.argtypes = [C.c_void_p]
        L.qs_gather_steps.restype = C.c_uint64
        L.qs_gather_close.argtypes = [C.c_void_p]
        handle = (C.c_ubyte * 64)()
        sim._check(L.qs_gather_create(sim.h, self.world, self.rank, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        buf = (C.c_ubyte * (64 * self.world))(*allh.cpu().tolist())
        sim._check(L.qs_gather_connect(sim.h, buf))
        dist.barrier(group=self.group)
        L.qs_gather_row_stride.argtypes = [C.c_void_p]
        stride = int(L.qs_gather_row_stride(sim.h))  # rows padded to 128 bytes: whole-line stores over NVLink
        self.gathered = [torch.as_tensor(_DevArray(L.qs_gather_buffer(sim.h, k), (self.world * self.N, stride)), device=self.dev)[:, :self.D]
                         for k in range(2)]

    def describe(self) -> str:
        if self.mode == 'p2p':
            return ('fused: each warp of the step kernel stores its observation row into the gathered tensor of every rank through '
                    'peer-mapped memory (CUDA IPC over NVLink), per-rank completion flags raised by the last warp of the launch; '
                    'wait() is a flag-poll kernel on the consumer stream')
        return 'NCCL all_gather_into_tensor on its own stream, double-buffered observation rows (overlaps the next step)'

    # ------------------------------------------------------------------ stepping
    def step_autoreset(self, ctrl, opt):
        k = self.t & 1
        main = torch.cuda.current_stream(self.dev)
        if self._rel_pending[k]:
            main.wait_event(self._ev_rel[k])  # explicit back-pressure: the consumer of step t - 2 has released this buffer
            self._rel_pending[k] = False
        if self.mode == 'nccl':
            self.sim.step_autoreset(ctrl, opt, obs_out=self.local[k])
            self._ev_step[k].record(main)
            with torch.cuda.stream(self._coll):
                self._coll.wait_event(self._ev_step[k])
                dist.all_gather_into_tensor(self.gathered[k], self.local[k], group=self.group)
                self._ev_gath[k].record(self._coll)
        else:
            self.sim.step_autoreset(ctrl, opt)  # rows go to every rank's gathered[k] from inside the kernel
            self._ev_step[k].record(main)
        self.t += 1

    def wait(self) -> torch.Tensor:
        """Make the CURRENT stream wait until the gathered tensor of the last step is complete on this rank; returns it."""
        k = (self.t - 1) & 1
        cur = torch.cuda.current_stream(self.dev)
        if self.mode == 'nccl':
            cur.wait_event(self._ev_gath[k])
        else:
            cur.wait_event(self._ev_step[k])  # own launch enqueued ...
            self.sim._check(self.sim.L.qs_gather_wait(self.sim.h, C.c_uint64(self.t), C.c_void_p(cur.cuda_stream)))  # ... and every peer's done
        return self.gathered[k]

    def release(self):
        """The consumer on the current stream is done with the tensor returned by the last `wait()`."""
        k = (self.t - 1) & 1
        self._ev_rel[k].record(torch.cuda.current_stream(self.dev))
        self._rel_pending[k] = True

    def close(self):
        torch.cuda.synchronize(self.dev)
        if self.mode == 'p2p':
            dist.barrier(group=self.group)
            self.sim.L.qs_gather_close(self.sim.h)


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that pinned host buffers allocated afterwards
    are first-touched in memory local to that GPU's PCIe root: with one process per GPU on a multi-socket host the zero-copy
    observation rows of `qs_step_host` otherwise cross the inter-socket link for half of the ranks.  Returns a description of what
    was done ('' if the topology could not be read; never raises)."""
    import os
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bus = None
        if all(hasattr(prop, k) for k in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            bus = f'{int(prop.pci_domain_id):04x}:{int(prop.pci_bus_id):02x}:{int(prop.pci_device_id):02x}.0'
        if bus is None:
            import subprocess
            out = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(device_index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bus = out.splitlines()[0].strip() if out else None
        if not bus:
            return ''
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:  # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        node_path = f'/sys/bus/pci/devices/{bus}/numa_node'
        node = int(open(node_path).read().strip())
        if node < 0:
            return ''
        cpus = open(f'/sys/devices/system/node/node{node}/cpulist').read().strip()
        ids = set()
        for part in cpus.split(','):
            lo, _, hi = part.partition('-')
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = ids & os.sched_getaffinity(0)
        if not allowed:
            return ''
        os.sched_setaffinity(0, allowed)
        return f'gpu {device_index} ({bus}) -> NUMA node {node}, {len(allowed)} cpus'
    except Exception:  # noqa: BLE001
        return ''
