"""Aggregate ncu warp-stall samples per CUDA source line from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > out.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
by_line = collections.Counter(); src_text = {}
def num(s):
    try: return int(s)
    except Exception: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] != '':
        try: ln = int(r[0])
        except Exception: continue
        by_line[(cur_file, ln)] += num(r[6]); src_text[(cur_file, ln)] = r[1]
tot = sum(by_line.values()); print('total samples', tot)
for (f, ln), s in by_line.most_common(top):
    print(f"{f}:{ln}  {s} ({100*s/tot:.1f}%)  {src_text[(f,ln)].strip()[:110]}")
