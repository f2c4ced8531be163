"""Wall clock of the drop-in executables against what they replace, on files in /dev/shm (the record kept in profiles/):
  sam2pairs | sort    reference binary (T = 8) piped through the driver's GNU sort (microcket:479-480)
  sam2pairs ... sorted    ours, sorted on the GPU (no sort process)
  sam2pairs | sort    ours, plain, piped through the same sort
  pairs2bins          ours on the sorted .pairs at the nine default resolutions (no reference program exists: juicer is a jar)
usage: python tools/cli_wallclock.py [read groups = 10000000] [mode = flash]"""
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
import bench as B  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "flash"
BIN = os.path.join(ROOT, "microcket_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref", "sam2pairs")
SORT = "LANG=C sort -k2,2d -k4,4d -k3,3n -k5,5n --parallel=8 -S 50%"
res = {"read_groups": n, "mode": mode, "runs": []}
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    src = os.path.join(td, "in.sam")
    nb_total = 0
    with open(src, "wb") as f:
        for k in range(0, n, 2_000_000):
            buf, nb = mk.synth_device(torch, 91, mode, "hg38", k, min(2_000_000, n - k), opts=mk.synth_opts(dup_per_1024=128, dup_universe=n))
            f.write(buf[:nb].cpu().numpy().tobytes()); nb_total += nb
    res["sam_GB"] = nb_total / 1e9
    torch.cuda.empty_cache()

    def run(name, cmd):
        t0 = time.perf_counter()
        subprocess.run(cmd, shell=True, check=True, cwd=td, executable="/bin/bash")
        dt = time.perf_counter() - t0
        res["runs"].append({"what": name, "seconds": round(dt, 2), "M_groups_per_s": round(n / dt / 1e6, 3)})
        return dt

    if os.path.exists(REF):
        run("reference sam2pairs (T=8) | sort", f"{REF} in.sam {mode} R 8 0.5 10 0 2>/dev/null | {SORT} > ref.pairs")
    run("ours sam2pairs (warm-up, not counted)", f"{BIN}/sam2pairs in.sam {mode} W 8 0.5 10 0 2>/dev/null > /dev/null")
    run("ours sam2pairs | sort", f"{BIN}/sam2pairs in.sam {mode} P 8 0.5 10 0 2>/dev/null | {SORT} > ours_piped.pairs")
    run("ours sam2pairs ... sorted (GPU sort, no sort process)", f"{BIN}/sam2pairs in.sam {mode} S 8 0.5 10 0 sorted 2>/dev/null > ours_sorted.pairs")
    same = [subprocess.run(f"cmp -s {a} {b}", shell=True, cwd=td).returncode == 0
            for a, b in (("ours_piped.pairs", "ours_sorted.pairs"),) + ((("ref.pairs", "ours_sorted.pairs"),) if os.path.exists(REF) else ())]
    res["outputs_identical"] = all(same)
    info = os.path.join(td, "hg38.info")
    open(info, "w").write("".join(f"{a}\t{b}\n" for a, b in zip(B.HG38, B.HG38_LEN)))
    run("ours pairs2bins, nine default resolutions", f"{BIN}/pairs2bins -r {','.join(map(str, B.DEFAULT_RES))} ours_sorted.pairs C hg38.info 2>/dev/null")
    res["pairs_lines"] = int(subprocess.run("wc -l < ours_sorted.pairs", shell=True, cwd=td, capture_output=True, text=True).stdout)
print(json.dumps(res))
