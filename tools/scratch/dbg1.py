import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import microcket_b200 as mk, oracle_lib
orc = oracle_lib.load()
sam = open("/root/repo/tests/golden/appB_unc.sam","rb").read()
op, osam, ost = orc.sam2pairs(sam, "unc", threads=2)
s = mk.Sam2Pairs(mk.S2PConfig(mode="unc", threads=2, write_sam=True))
p, so, st = s.run(sam)
print("pairs equal", p == op, "len", len(so), len(osam))
i = 0
for k,(x,y) in enumerate(zip(so, osam)):
    if x != y:
        print("first diff at", k); break
# show zero runs
import re
for m in re.finditer(rb"\x00+", so):
    print("zeros", m.start(), m.end()-m.start(), "line start?", m.start()==0 or so[m.start()-1:m.start()]==b"\n")
ls = osam.split(b"\n")
off = 0
for l in ls[:12]:
    print(off, len(l)+1, off % 16); off += len(l)+1
