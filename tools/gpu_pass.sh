#!/bin/bash
# tests of the sam2pairs path, a 30 M-group A/B bench line, then the full profile pass (tools/profile_round.sh)
TAG=${1:-rXX}
timeout 600 python -m pytest tests/test_gpu_s2p.py -x -q 2>&1 | tail -3
python bench.py --groups 30000000 --no-cpu --no-e2e > gpurun_out/${TAG}_bench30.json 2> gpurun_out/${TAG}_bench30.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench30.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()})
PY
tools/profile_round.sh $TAG 2>&1 | tail -3
