"""Aggregate ncu per-source-line stall samples of k_solve by code region, barrier stalls apart
   (waiting warps) from the rest (working warps).
   usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > out.csv ; python tools/ncu_regions.py out.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur = None
nb = collections.Counter(); tot = collections.Counter(); bar = collections.Counter(); inst = collections.Counter()
def num(s):
    try: return int(s)
    except Exception: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if not re.fullmatch(r'\d+', r[0] or ''): continue
    k = (cur, int(r[0])); s = num(r[6]); b = 0
    for i in range(32, min(49, len(r))):
        if hdr[i] == 'stall_barrier': b = num(r[i])
    tot[k] += s; nb[k] += s - b; bar[k] += b; inst[k] += num(r[7])
# function ranges from the sources
import os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'landing_controller_b200', 'csrc')
fn_of = {}
for f in ('sweeps.cuh', 'solver_dev.cuh', 'srb_knot.cuh'):
    name = f
    marks = []
    for i, l in enumerate(open(os.path.join(root, f)), 1):
        m = re.match(r'^(?:template.*>\s*)?__device__.*?\b(\w+)\(', l) or re.match(r'^__global__.*?\b(\w+)\(', l)
        if m: name = m.group(1)
        mm = re.search(r'//\s*(S\d)\.', l)
        if mm and name == 'backward_sweep': marks.append(1)
        if name == 'backward_sweep':
            tag = re.search(r'^\s*//\s*(S\d)\.', l)
            if tag: cur_tag = tag.group(1)
            elif 'backward_sweep(' in l: cur_tag = 'init'
            fn_of[(f, i)] = 'backward_sweep:' + cur_tag
        else:
            fn_of[(f, i)] = name
T = sum(tot.values())
reg_nb = collections.Counter(); reg_b = collections.Counter(); reg_i = collections.Counter()
for k in tot:
    r = fn_of.get(k, k[0])
    reg_nb[r] += nb[k]; reg_b[r] += bar[k]; reg_i[r] += inst[k]
TI = sum(reg_i.values())
print('samples %d (barrier %.1f%%), warp instructions %d' % (T, 100.0 * sum(bar.values()) / T, TI))
for k, v in sorted(reg_nb.items(), key=lambda kv: -(kv[1] + reg_b[kv[0]])):
    if v + reg_b[k] < 0.002 * T: continue
    print('%-28s working %6.2f%%  at-barrier %6.2f%%  inst %6.2f%%' % (k, 100 * v / T, 100 * reg_b[k] / T, 100 * reg_i[k] / TI))
