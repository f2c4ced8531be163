"""Turn an .ncu-rep (ncu --set full) into the small CSV kept under profiles/ and, optionally, profiles/traffic.json.

usage: python tools/ncu_summary.py gpurun_out/r01c_s2p.ncu-rep profiles/r01c_s2p_ncu_summary.csv [--traffic profiles/traffic.json --window-mb 1024]
"""
import csv
import io
import json
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])
    if "--traffic" in sys.argv:
        tp = sys.argv[sys.argv.index("--traffic") + 1]
        wmb = int(sys.argv[sys.argv.index("--window-mb") + 1]) if "--window-mb" in sys.argv else None
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        t = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes, from ncu --set full --clock-control none on one window "
                      "of the bench workload (%s); first captured launch of each kernel" % out, "window_mb": wmb}
        for r in body:
            name = r[ik].split("(")[0].replace("void ", "").split("<")[0]
            if name in t:
                continue
            t[name] = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        json.dump(t, open(tp, "w"), indent=1)


if __name__ == "__main__":
    main()
