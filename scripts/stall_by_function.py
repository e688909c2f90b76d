"""Per device function: executed instructions and stall samples split by reason (ncu source page x nvdisasm line table).
usage: python scripts/stall_by_function.py <tag> [kernel-substring]"""
import bisect, collections, csv, re, subprocess, sys, tempfile, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
KERNEL = sys.argv[2] if len(sys.argv) > 2 else 'env_kernelIfLi16ELi3ELi0ELi8101E'
rep = ROOT / 'gpurun_out' / f'prof_{tag}.ncu-rep'
tmp = Path(tempfile.mkdtemp())
subprocess.run(['cuobjdump', '-xelf', 'all', str(ROOT / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so')], cwd=tmp, capture_output=True)
dis = None
for cubin in sorted(tmp.glob('*.cubin')):
    d_ = subprocess.run(['nvdisasm', '-g', '-c', str(cubin)], capture_output=True, text=True).stdout.split('\n')
    hit = [i for i, l in enumerate(d_) if l.startswith('.text.') and KERNEL in l and l.rstrip().endswith(':')]
    if hit:
        dis, start = d_, hit[0]
        break
insts, cur = [], ('?', 0)
for l in dis[start + 1:]:
    if l.startswith('//---------------------'):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*?;', l):
        insts.append(cur)
src = subprocess.run(['ncu', '-i', str(rep), '--page', 'source', '--csv'], capture_output=True, text=True, check=True).stdout
srows = list(csv.reader(src.splitlines()))
shdr, sdata = srows[1], srows[2:]
ci = shdr.index('Instructions Executed')
reasons = ['stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_branch_resolving', 'stall_not_selected', 'stall_membar', 'stall_dispatch', 'stall_math', 'stall_selected', 'stall_barrier', 'stall_lg']
ridx = [shdr.index(r) for r in reasons]
def funcs(path):
    res = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:QS_DEV|QS_NOINLINE)\s+(?:static\s+)?[\w:<>\s\*&]+?\s(\w+)\(', l)
        if m and not l.strip().startswith('//'):
            res.append((i, m.group(1)))
    return res
fmap = {f: funcs(ROOT / 'gym_quadruped_b200' / 'csrc' / f) for f in ('qs_env.cuh', 'qs_math.cuh')}
dyn = collections.Counter(); st = collections.defaultdict(lambda: collections.Counter())
by_line = collections.defaultdict(lambda: collections.Counter())
for k in range(min(len(insts), len(sdata))):
    f, ln = insts[k]
    try: ie = int(sdata[k][ci])
    except ValueError: ie = 0
    if f in fmap:
        fl = fmap[f]; idx = bisect.bisect_right([x[0] for x in fl], ln) - 1
        name = f.split('.')[0][3:] + ':' + (fl[idx][1] if idx >= 0 else '?')
    else:
        name = f
    dyn[name] += ie
    for r, j in zip(reasons, ridx):
        try: v = int(sdata[k][j])
        except ValueError: v = 0
        st[name][r] += v
        if f == 'qs_kernel.cuh': by_line[ln][r] += v
tot = collections.Counter()
for n in st: tot.update(st[n])
print('reason totals:', {r: tot[r] for r in reasons})
print('%-30s %8s %7s | ' % ('function', 'dyn/env', 'samples') + ' '.join('%8s' % r[6:14] for r in reasons[:9]))
for name, v in sorted(dyn.items(), key=lambda kv: -sum(st[kv[0]].values()))[:32]:
    s = st[name]
    print('%-30s %8.0f %7d | ' % (name, v / 4096, sum(s.values())) + ' '.join('%8d' % s[r] for r in reasons[:9]))
print('\nqs_kernel.cuh lines with most samples:')
for ln, c in sorted(by_line.items(), key=lambda kv: -sum(kv[1].values()))[:25]:
    print(ln, sum(c.values()), {k[6:]: v for k, v in c.items() if v > 5})
shutil.rmtree(tmp)
