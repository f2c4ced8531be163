"""Per-source-line totals of one kernel from an .ncu-rep captured with --import-source on (development aid).
usage: python tools/ncu_lines.py rep.ncu-rep <kernel regex> [N]   -> top N source lines by stall samples and by instructions"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; smp = collections.Counter(); ins = collections.Counter(); thr = collections.Counter(); src = {}; why = collections.defaultdict(collections.Counter)
conf = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)                      # first 'Source' = the CUDA line, second = SASS
        key = d["Line No"]
        try:
            s = int(d["# Samples"] or 0); i = int(d["Instructions Executed"] or 0)
        except ValueError:
            continue
        smp[key] += s; ins[key] += i; thr[key] += int(d["Thread Instructions Executed"] or 0); src[key] = d["Source"].strip()[:120]
        conf[key] += int(d.get("L1 Wavefronts Shared Excessive") or 0)
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                try: why[key][k[6:]] += int(d[k] or 0)
                except ValueError: pass
ts, ti = sum(smp.values()) or 1, sum(ins.values()) or 1
print("total samples", ts, "warp instructions", ti)
print("--- by samples")
for k, v in smp.most_common(N):
    top = ",".join(f"{a}:{b}" for a, b in why[k].most_common(3))
    print(f"{100*v/ts:5.1f}% smp {100*ins[k]/ti:5.1f}% ins  thr/ins {thr[k]/max(ins[k],1):4.1f} conf {conf[k]:8d} [{top}] {k}: {src[k]}")
print("--- by instructions")
for k, v in ins.most_common(N):
    print(f"{100*smp[k]/ts:5.1f}% smp {100*v/ti:5.1f}% ins  thr/ins {thr[k]/max(v,1):4.1f} {k}: {src[k]}")
