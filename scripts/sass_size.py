"""Static SASS size and opcode histogram per kernel of libqstep.so (cuobjdump -sass); used for profiles/r02_sass_*.txt."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1] if len(sys.argv) > 1 else Path(__file__).resolve().parents[1] / 'gym_quadruped_b200' / 'csrc' / 'libqstep.so')
want = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True).stdout
fn, sizes, ops = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = m.group(1); sizes[fn] = 0; ops[fn] = collections.Counter(); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        sizes[fn] += 1
        ops[fn][m.group(2).split('.')[0]] += 1
for fn, n in sizes.items():
    if want and want not in fn:
        continue
    c = ops[fn]
    print(f'{fn}: {n} instructions = {n * 16 / 1024:.0f} KB; STL {c["STL"]} LDL {c["LDL"]} UBLKCP {c["UBLKCP"]} SYNCS {c["SYNCS"]} SHFL {c["SHFL"]} MUFU {c["MUFU"]} BSSY {c["BSSY"]} '
          f'FFMA {c["FFMA"]} FMUL {c["FMUL"]} FADD {c["FADD"]} LDS {c["LDS"]} STS {c["STS"]} ACQBULK {c["ACQBULK"]}')
    if want:
        for op, k in c.most_common(40):
            print(f'    {op:12s} {k}')
