#!/bin/bash
# GPU parity tests + one 30 M-group bench line (no CPU / e2e legs): the A/B loop of a kernel change
TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --groups 30000000 --no-cpu --no-e2e > gpurun_out/${TAG}_bench30.json 2> gpurun_out/${TAG}_bench30.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench30.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()})
PY
