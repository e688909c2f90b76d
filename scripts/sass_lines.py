"""Diagnostic: dump the SASS of the step kernel attributed to a source-line range, with executed counts and stall samples.
    python scripts/sass_lines.py <tag> <file> <first> <last> [lib]"""
import csv, re, subprocess, sys, tempfile, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
lib = sys.argv[5] if len(sys.argv) > 5 else str(ROOT / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so')
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
ci, si, so = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
cols = {k: hdr.index(k) for k in hdr if k.startswith('stall_') and '(' not in k}
idx = [k for k in range(min(len(insts), len(data))) if insts[k][0] == fname and lo <= insts[k][1] <= hi]
if idx:
    for k in range(min(idx), max(idx) + 1):
        r = data[k]
        st = ' '.join('%s=%s' % (n[6:], r[c]) for n, c in cols.items() if r[c] not in ('0', ''))
        print(f'{k:6d} {insts[k][0][:10]}:{insts[k][1]:<5d} x{int(r[ci] or 0)/4096:6.2f} s{r[si]:>4s}  {r[so].strip()[:70]:70s} | {st}')
shutil.rmtree(tmp)
