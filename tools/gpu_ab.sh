#!/bin/bash
# A/B loop of a sam2pairs kernel change: parity tests of the path, a 30 M-group bench line, and (NCU=1) one `ncu --set full`
# capture with source of the per-line kernels on the first 2040 MiB window.   usage: tools/gpu_ab.sh <tag>
TAG=${1:-q}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_s2p.py -x -q 2>&1 | tail -3
python bench.py --groups 30000000 --no-cpu --no-e2e > $O/${TAG}_bench30.json 2> $O/${TAG}_bench30.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench30.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()}, d["roofline"].get("pairs_stage", {}).get("ms_per_step"))
PY
if [ -n "$NCU" ]; then
  SMALL="--groups 6000000 --steps 2 --warmup 1 --no-cpu --no-e2e"
  ncu --set full --clock-control none --import-source on -k regex:'k_scan_chunks|k_chunk_prefix|k_chunk_compact|k_parse|k_group|k_emit' \
      -s 21 -c 7 -f -o $O/${TAG}_s2p python bench.py $SMALL > $O/${TAG}_ncu_full.log 2>&1
  ls -la $O/${TAG}_s2p.ncu-rep
fi
