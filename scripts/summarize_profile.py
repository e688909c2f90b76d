"""Turns gpurun_out/prof_<tag>.ncu-rep (+ launches_<tag>.csv) into the committed summaries under profiles/.

    python scripts/summarize_profile.py <tag> <round-label>

Needs ncu, cuobjdump and nvdisasm (no GPU).  Writes:
  profiles/<label>_step_kernel_metrics.csv   selected raw metrics of the step kernel (ncu --set full)
  profiles/<label>_function_attribution.txt  executed instructions / stall samples per device function (SASS -> source via -lineinfo)
  profiles/<label>_launches.csv              the ncu launch list (gpu__time_duration.sum per kernel launch)
  profiles/traffic.json                      DRAM bytes per launch of the step kernel (read by bench.py: roofline.traffic)
"""
import bisect
import collections
import csv
import json
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, label = sys.argv[1], sys.argv[2]
rep = ROOT / 'gpurun_out' / f'prof_{tag}.ncu-rep'
out = ROOT / 'profiles'
out.mkdir(exist_ok=True)
KERNEL = sys.argv[3] if len(sys.argv) > 3 else 'env_kernelIfLi16ELi3ELi0ELi8101E'

raw = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'sm__cycles_elapsed.max']
sel = [(h, units[i], data[i]) for i, h in enumerate(hdr) if h in keep or h.startswith('smsp__average_warps_issue_stalled')]
with open(out / f'{label}_step_kernel_metrics.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit', 'value'])
    w.writerows(sel)
vals = {h: (u, v) for h, u, v in sel}


def to_bytes(name):
    u, v = vals[name]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    return float(v) * scale


traffic = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
(out / 'traffic.json').write_text(json.dumps({'step_kernel_dram_bytes_per_launch': traffic, 'source': f'{label}_step_kernel_metrics.csv',
                                             'note': 'ncu --set full, one launch of env_kernel<float,16,3,MODE_STEP>, 4096 envs'}))

# ---- SASS -> source attribution
tmp = Path(tempfile.mkdtemp())
subprocess.run(['cuobjdump', '-xelf', 'all', str(ROOT / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so')], cwd=tmp, capture_output=True)
dis, start = None, None
for cubin in sorted(tmp.glob('*.cubin')):  # one cubin per translation unit (qs_inst_*.cu): find the one that holds the kernel
    d_ = subprocess.run(['nvdisasm', '-g', '-c', str(cubin)], capture_output=True, text=True).stdout.split('\n')
    hit = [i for i, l in enumerate(d_) if l.startswith('.text.') and KERNEL in l and l.rstrip().endswith(':')]
    if hit:
        dis, start = d_, hit[0]
        break
assert dis is not None, f'{KERNEL} not found in libqstep.so'
insts, cur = [], ('?', 0)
for l in dis[start + 1:]:
    if l.startswith('//---------------------'):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*?;', l):
        insts.append(cur)
src = subprocess.run(['ncu', '-i', str(rep), '--page', 'source', '--csv'], capture_output=True, text=True, check=True).stdout
srows = list(csv.reader(src.splitlines()))
shdr, sdata = srows[1], srows[2:]
ci, si = shdr.index('Instructions Executed'), shdr.index('# Samples')


def funcs(path):
    res = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:QS_DEV|QS_NOINLINE)\s+(?:static\s+)?[\w:<>\s\*&]+?\s(\w+)\(', l)
        if m and not l.strip().startswith('//'):
            res.append((i, m.group(1)))
    return res


fmap = {f: funcs(ROOT / 'gym_quadruped_b200' / 'csrc' / f) for f in ('qs_env.cuh', 'qs_math.cuh')}
dyn, smp, stat = collections.Counter(), collections.Counter(), collections.Counter()
n = min(len(insts), len(sdata))
for k in range(n):
    f, ln = insts[k]
    try:
        ie, sm = int(sdata[k][ci]), int(sdata[k][si])
    except ValueError:
        ie = sm = 0
    if f in fmap:
        fl = fmap[f]
        idx = bisect.bisect_right([x[0] for x in fl], ln) - 1
        name = f.split('.')[0][3:] + ':' + (fl[idx][1] if idx >= 0 else '?')
    else:
        name = f
    dyn[name] += ie; smp[name] += sm; stat[name] += 1
tot, stot = sum(dyn.values()), sum(smp.values())
nwarps = 4096
with open(out / f'{label}_function_attribution.txt', 'w') as f:
    f.write(f'# step kernel {KERNEL}: static SASS instructions {len(insts)}, executed warp-instructions {tot} '
            f'({tot / nwarps:.0f} per env-step), stall samples {stot}\n')
    f.write('# "math:umulhi" collects the small inlined helpers of qs_math.cuh (dot3, cross3, Num<>::..., reductions)\n')
    f.write('%-34s %8s %10s %6s %8s %6s\n' % ('function', 'static', 'dyn/env', '%', 'samples', '%'))
    for name, v in sorted(dyn.items(), key=lambda kv: -kv[1]):
        f.write('%-34s %8d %10.0f %5.1f%% %8d %5.1f%%\n' % (name, stat[name], v / nwarps, 100 * v / max(1, tot), smp[name], 100 * smp[name] / max(1, stot)))
shutil.rmtree(tmp)
lp = ROOT / 'gpurun_out' / f'launches_{tag}.csv'
if lp.exists():
    shutil.copy(lp, out / f'{label}_launches.csv')
print('time us', vals['gpu__time_duration.sum'], 'inst', vals['smsp__inst_executed.sum'], 'dram bytes/launch', traffic)
