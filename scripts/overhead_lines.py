import csv, re, subprocess, tempfile, shutil, collections, sys
from pathlib import Path
ROOT=Path('/root/repo'); KERNEL='env_kernelIfLi16ELi3ELi0E'; tag=sys.argv[1]
tmp=Path(tempfile.mkdtemp())
subprocess.run(['cuobjdump','-xelf','all',str(ROOT/'gym_quadruped_b200/csrc/libqstep.so')],cwd=tmp,capture_output=True)
cubin=next(tmp.glob('*.cubin'))
dis=subprocess.run(['nvdisasm','-g','-c',str(cubin)],capture_output=True,text=True).stdout.split('\n')
start=[i for i,l in enumerate(dis) if l.startswith('.text.') and KERNEL in l and l.rstrip().endswith(':')][0]
insts=[];cur=('?',0)
for l in dis[start+1:]:
    if l.startswith('//---------------------'): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m2=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);',l)
    if m2: insts.append((cur,m2.group(1)))
src=subprocess.run(['ncu','-i',str(ROOT/f'gpurun_out/prof_{tag}.ncu-rep'),'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=rows[2:]
ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples')
FP={'FFMA','FMUL','FADD','MUFU','FSEL','FSETP','FMNMX','LDS','STS','SHFL','LDG','STG','DADD','DMUL','DFMA','F2F','LD','ST','ATOMG','RED','LDL','STL'}
byline=collections.Counter(); allline=collections.Counter(); tot=0; ops=collections.Counter(); smp=collections.Counter()
for k in range(min(len(insts),len(data))):
    (f,ln),txt=insts[k]
    op=re.sub(r'^@!?U?P\d+\s+','',txt).split()[0].split('.')[0]
    e=int(data[k][ci] or 0); tot+=e; ops[op]+=e; allline[(f,ln)]+=e; smp[(f,ln)]+=int(data[k][si] or 0)
    if op not in FP: byline[(f,ln)]+=e
print('total/env',round(tot/4096), 'non-FP/mem share', round(sum(byline.values())/tot,3))
print({o:round(c/4096) for o,c in ops.most_common(22)})
text={f:open(ROOT/'gym_quadruped_b200/csrc'/f).read().split('\n') for f in ('qs_env.cuh','qs_math.cuh','qstep.cu')}
for (f,ln),c in byline.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 30):
    t=text[f][ln-1].strip()[:95] if f in text and ln-1 < len(text[f]) else ''
    print(f'{c/4096:6.1f} /{allline[(f,ln)]/4096:6.1f} s{smp[(f,ln)]:4d} {f}:{ln} {t}')
shutil.rmtree(tmp)
