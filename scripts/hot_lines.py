"""Diagnostic: rank source lines of the step kernel by ncu stall samples (needs gpurun_out/prof_<tag>.ncu-rep and the matching libqstep.so).
    python scripts/hot_lines.py <tag> [top]"""
import collections, csv, re, subprocess, sys, tempfile, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
lib = sys.argv[3] if len(sys.argv) > 3 else str(ROOT / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so')
KERNEL = 'env_kernelIfLi16ELi3ELi0E'
tmp = Path(tempfile.mkdtemp())
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
cubin = next(tmp.glob('*.cubin'))
dis = subprocess.run(['nvdisasm', '-g', '-c', str(cubin)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and KERNEL in l and l.rstrip().endswith(':')][0]
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
src = subprocess.run(['ncu', '-i', str(ROOT / 'gpurun_out' / f'prof_{tag}.ncu-rep'), '--page', 'source', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
smp, dyn = collections.Counter(), collections.Counter()
for k in range(min(len(insts), len(data))):
    try:
        smp[insts[k]] += int(data[k][si]); dyn[insts[k]] += int(data[k][ci])
    except ValueError:
        pass
tot = sum(smp.values())
text = {f: open(ROOT / 'gym_quadruped_b200' / 'csrc' / f).read().split('\n') for f in ('qs_env.cuh', 'qs_math.cuh', 'qstep.cu')}
print('total samples', tot)
for (f, ln), s in smp.most_common(top):
    t = text[f][ln - 1].strip()[:110] if f in text and ln - 1 < len(text[f]) else ''
    print(f'{s:5d} {100*s/tot:4.1f}%  dyn/env {dyn[(f, ln)]/4096:7.1f}  {f}:{ln}  {t}')
shutil.rmtree(tmp)
