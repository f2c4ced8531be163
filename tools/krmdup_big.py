"""Drop-in krmdup CLI against the reference binary on a large FASTQ (wall clock, files in /dev/shm): the record kept in
profiles/.   usage: python tools/krmdup_big.py [pairs = 10000000]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import microcket_b200 as mk  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ours = os.path.join(ROOT, "microcket_b200", "bin", "krmdup")
ref = os.path.join(ROOT, "oracle", "_ref", "krmdup")
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    src = os.path.join(td, "in.fq")
    nb_total = 0
    with open(src, "wb") as f:
        for k in range(0, n_pairs, 2_000_000):
            buf, nb = mk.synth_device(torch, 77, "fastq", "hg38", k, min(2_000_000, n_pairs - k))
            f.write(buf[:nb].cpu().numpy().tobytes()); nb_total += nb
    res = {"pairs": n_pairs, "fastq_GB": nb_total / 1e9, "runs": []}
    outs = {}
    for name, exe in (("gpu (cold)", ours), ("gpu", ours), ("gpu", ours), ("reference", ref)):
        if not os.path.exists(exe):
            continue
        pre = os.path.join(td, "o")
        for e in ("read1.fq", "read2.fq", "log"):
            if os.path.exists(f"{pre}.{e}"):
                os.remove(f"{pre}.{e}")
        t0 = time.perf_counter()
        subprocess.run([exe, "-i", src, "-o", pre], check=True)
        dt = time.perf_counter() - t0
        import hashlib
        h = hashlib.sha1()
        for e in ("read1.fq", "read2.fq", "log"):
            h.update(open(f"{pre}.{e}", "rb").read())
        outs[name.split()[0]] = h.hexdigest()
        res["runs"].append({"program": name, "seconds": round(dt, 3), "M_pairs_per_s": round(n_pairs / dt / 1e6, 3), "GB_per_s": round(nb_total / dt / 1e9, 3)})
    res["outputs_identical"] = len(set(outs.values())) == 1
    print(json.dumps(res))
