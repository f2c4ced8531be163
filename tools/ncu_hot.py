"""Top source lines of a kernel from an .ncu-rep captured with --import-source on (development aid).
usage: python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] and hdr:
        d = dict(zip(hdr, r))
        try:
            out.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), float(d["Avg. Threads Executed"] or 0), cur, r[0], r[1].strip()[:110]))
        except ValueError:
            pass
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print("total samples", ts, "total warp instructions", ti)
print("--- by samples")
for s, i, t, f, ln, src in sorted(out, reverse=True)[:N]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins thr {t:4.1f} {f}:{ln}  {src}")
print("--- by instructions")
for s, i, t, f, ln, src in sorted(out, key=lambda o: -o[1])[:N]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins thr {t:4.1f} {f}:{ln}  {src}")
