#!/bin/bash
# One GPU-box pass that produces everything profiles/ cites: full-size bench line (+ reference arm), the other BASELINE
# configurations, the CUB A/B, the ncu launch list of the bench command at a size ncu can replay, and `ncu --set full`
# captures (with source) of the sam2pairs kernels, the sort and the histogram.
# usage: tools/profile_round.sh <tag>        (outputs under gpurun_out/<tag>_*)
set -x
TAG=${1:-rXX}
O=gpurun_out
python bench.py > $O/${TAG}_bench_full.json 2> $O/${TAG}_bench_full.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
for c in "--config unc" "--config multires" "--sam"; do
  t=$(echo $c | tr -d " -")
  python bench.py $c --groups 30000000 --no-e2e --cpu-groups 2000000 > $O/${TAG}_bench_$t.json 2> $O/${TAG}_bench_$t.err
done
[ -x tools/scratch/cub_ab ] && tools/scratch/cub_ab 97800000 69 5 > $O/${TAG}_cub_ab.json
SMALL="--groups 6000000 --steps 2 --warmup 1 --no-cpu --no-e2e"
python bench.py $SMALL > $O/${TAG}_bench_small.json 2> $O/${TAG}_bench_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py $SMALL > $O/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scan_chunks|k_chunk_prefix|k_chunk_compact|k_parse|k_group|k_emit' \
    -s 21 -c 7 -f -o $O/${TAG}_s2p python bench.py $SMALL > $O/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_radix_hist|k_uniq_cells|k_pack_keys' \
    -s 12 -c 5 -f -o $O/${TAG}_sort python bench.py $SMALL > $O/${TAG}_ncu_sort.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_hist_add|k_hist_coo' \
    -s 6 -c 6 -f -o $O/${TAG}_hist python bench.py --config multires $SMALL > $O/${TAG}_ncu_hist.log 2>&1
tail -c 600 $O/${TAG}_bench_full.json
