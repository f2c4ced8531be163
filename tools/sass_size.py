"""Static SASS size of one kernel by source line / sub-function (development aid; needs -lineinfo).
usage: python tools/sass_size.py microcket_b200/csrc/_obj/s2p.o k_ft_strip [N]"""
import re, subprocess, sys, os, tempfile
obj, kern = sys.argv[1], sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur = None; agg = {}; fn = {}; infn = False; sub = kern
for l in out.splitlines():
    if l.startswith("\t.section\t.text."): infn = kern in l; sub = kern; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\$?.*\$(_Z\w+):', l.strip())
    if m: sub = m.group(1)
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+[A-Z@]', l):
        agg[cur] = agg.get(cur, 0) + 1; fn[sub] = fn.get(sub, 0) + 1
print("total", sum(agg.values()))
for k, v in sorted(fn.items(), key=lambda kv: -kv[1]): print(f"{v:6d} {k[:100]}")
print("--- lines")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:N]: print(v, k)
