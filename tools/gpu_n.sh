#!/bin/bash
# One N-GPU bench line under torchrun (as the driver launches it) + its summary.   usage: tools/gpu_n.sh <N> <tag> [bench args]
N=$1; TAG=$2; shift 2
O=gpurun_out
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" \
    > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err
grep "bench\]" $O/${TAG}_n$N.err | cut -c1-400
python - <<PY
import json
d = json.loads(open("$O/${TAG}_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"].get("exchange_stage"), d.get("e2e") and d["e2e"].get("value"))
print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["pairs_stage"]["ms_per_step"])
PY
