"""Static SASS instruction count per source line of a cubin/object (code-size diet of k_solve).
   usage: python tools/sass_lines.py landing_controller_b200/csrc/build/solver.o [top]"""
import re, collections, subprocess, sys, tempfile, os, glob
obj = os.path.abspath(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = glob.glob(d + "/*.cubin")[0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; fn = None; byfn = collections.Counter()
for l in txt.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur: cnt[cur] += 1
tot = sum(cnt.values()); print("total instructions", tot, "=", tot * 16 // 1024, "KB")
byfile = collections.Counter()
for (f, ln), c in cnt.items(): byfile[f] += c
print(byfile.most_common(8))
for (f, ln), c in cnt.most_common(top): print("%-22s %5d  %5d instr  %5.1f KB" % (f, ln, c, c * 16 / 1024))
