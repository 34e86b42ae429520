"""Aggregate ncu source-page CSV (ncu -i X --page source --csv --print-source cuda,sass) per device function / line."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
niter = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cur=None; agg=collections.Counter(); smp=collections.Counter(); lines=collections.Counter(); txt={}
def num(s):
    try: return int(s)
    except Exception: return 0
def func_ranges(path):
    out=[]
    for i,l in enumerate(open(path),1):
        m=re.match(r'^(?:template <[^>]*>\s*)?__device__ .*?(\w+)\(', l)
        if m: out.append((i,m.group(1)))
    return out
fr={f:func_ranges('landing_controller_b200/csrc/'+f) for f in ('sweeps.cuh','solver_dev.cuh')}
def fn(f,ln):
    r=fr.get(f)
    if not r: return f
    name=f+':?'
    for s,n in r:
        if s<=ln: name=n
    return name
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    try: ln=int(r[0])
    except Exception: continue
    k=fn(cur,ln)
    agg[k]+=num(r[7]); smp[k]+=num(r[6]); lines[(cur,ln)]+=num(r[7]); txt[(cur,ln)]=r[1]
ti=sum(agg.values()); ts=sum(smp.values())
print('total warp instructions', ti, 'per unit', ti/niter)
for k,v in agg.most_common(30): print('%-28s inst %5.1f%%  (%8.0f /unit) samples %5.1f%%'%(k,100*v/ti,v/niter,100*smp[k]/max(ts,1)))
print()
for (f,ln),v in lines.most_common(30): print('%s:%d %5.1f%% %s'%(f,ln,100*v/ti,txt[(f,ln)].strip()[:100]))
