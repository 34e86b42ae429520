"""Aggregate ncu per-source-line metrics from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > out.csv
   usage: ncu_lines.py out.csv [top] [samples|inst]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else 'samples'
cur_file = None
hdr = None
by_line = collections.Counter(); src_text = {}; inst = collections.Counter()
stalls = collections.defaultdict(collections.Counter)
def num(s):
    try: return int(s)
    except Exception: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if r[0] != '':
        try: ln = int(r[0])
        except Exception: continue
        k = (cur_file, ln)
        by_line[k] += num(r[6]); inst[k] += num(r[7]); src_text[k] = r[1]
        if hdr:
            for i in range(32, 49):
                if i < len(r) and num(r[i]): stalls[k][hdr[i]] += num(r[i])
tot = sum(by_line.values()); toti = sum(inst.values())
print('total samples', tot, 'total warp instructions', toti)
order = by_line if key == 'samples' else inst
for (f, ln), s in order.most_common(top):
    k = (f, ln)
    st = ' '.join('%s:%d' % (a.replace('stall_', ''), b) for a, b in stalls[k].most_common(3))
    print(f"{f}:{ln}  smp {by_line[k]} ({100*by_line[k]/max(tot,1):.1f}%) inst {inst[k]} ({100*inst[k]/max(toti,1):.1f}%) [{st}] {src_text[k].strip()[:90]}")
