import re,csv,collections,sys
"""Inclusive per-function profile of one kernel from an ncu source-page export.
usage: ncu_inclusive_profile.py <ncu --page source --csv export> <kernel name fragment> <nvdisasm -gi dump of the SAME cubin> <csrc dir of the SAME commit>
Every SASS instruction is charged to each function in its inline chain (nvdisasm -gi prints the chain)."""
prof, variant, dis, srcdir = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
src={f:open(srcdir+'/'+f).read().split('\n') for f in ('rt_core.cuh','render.cu')}
def func_of(f_,ln):
    if f_ not in src: return f_
    for k in range(min(ln,len(src[f_]))-1,-1,-1):
        m=re.match(r'^(?:MORT_HD(?:_NOINLINE)?|__global__|__device__|static|inline|template).*?\b([A-Za-z_0-9]+)\s*\(', src[f_][k])
        if m and not src[f_][k].startswith(' '):
            if m.group(1)=='__launch_bounds__': return 'mega_kernel'
            return m.group(1)
    return f_
lines=[]; chain=[]; fn=None; pending=[]
for l in open(dis):
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m=re.match(r'\s*\.text\.(\S+):', l)
    if m: fn=m.group(1); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and fn and variant in fn:
        if pending: chain=pending; pending=[]
        lines.append((int(m.group(1),16), m.group(2), list(chain)))
    elif m: pending=[]
lines.sort(key=lambda x:x[0])
rows=list(csv.reader(open(prof))); hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
print('instrs', len(lines), len(data))
incl=collections.defaultdict(lambda:[0,0,0]); tot=0; stot=0
extra=collections.defaultdict(lambda:[0,0,0])
for (addr,txt,ch),r in zip(lines,data):
    ie=float(r[ix['Instructions Executed']] or 0); te=float(r[ix['Thread Instructions Executed']] or 0); sm=float(r[ix['# Samples']] or 0)
    tot+=ie; stot+=sm
    fns=[]
    for (f_,ln) in ch:
        fn_=func_of(f_,ln)
        if fn_ not in fns: fns.append(fn_)
        if fn_=='quad_test':
            key='quad_test:interior' if ln>=216 else 'quad_test:plane'
            if key not in fns: fns.append(key)
    for fn_ in fns:
        a=incl[fn_]; a[0]+=ie; a[1]+=te; a[2]+=sm
print('total warp instr %.4g'%tot)
for k,a in sorted(incl.items(), key=lambda x:-x[1][0])[:45]:
    print(f'{100*a[0]/tot:5.1f}% inst {100*a[2]/stot:5.1f}% samp lanes {a[1]/max(a[0],1):5.1f}  {k}')
