"""krmdup (FASTQ path): our drop-in executable against the reference binary on the same file (development aid)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import microcket_b200 as mk

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
fq = b"".join(mk.synth_host(77, "fastq", "hg38", k, min(200_000, n_pairs - k)) for k in range(0, n_pairs, 200_000))
print(f"{n_pairs} pairs, {len(fq) / 1e9:.2f} GB FASTQ")
ours = os.path.join(ROOT, "microcket_b200", "bin", "krmdup")
ref = os.path.join(ROOT, "oracle", "_ref", "krmdup")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
    src = os.path.join(td, "in.fq")
    open(src, "wb").write(fq)
    out = {}
    for name, exe in (("gpu", ours), ("gpu", ours), ("reference", ref)):
        if not os.path.exists(exe):
            continue
        pre = os.path.join(td, name)
        for e in ("read1.fq", "read2.fq", "log"):
            if os.path.exists(f"{pre}.{e}"):
                os.remove(f"{pre}.{e}")
        t0 = time.perf_counter()
        subprocess.run([exe, "-i", src, "-o", pre], check=True)
        dt = time.perf_counter() - t0
        out[name] = open(pre + ".read1.fq", "rb").read() + open(pre + ".read2.fq", "rb").read() + open(pre + ".log", "rb").read()
        print(f"{name:9s} krmdup: {dt:.3f} s  {n_pairs / dt / 1e6:.2f} M pairs/s  {len(fq) / dt / 1e9:.2f} GB/s")
    if len(out) == 2:
        print("outputs identical:", out["gpu"] == out["reference"])
