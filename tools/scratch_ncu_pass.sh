O=gpurun_out; TAG=w1
SMALL="--groups 6000000 --steps 2 --warmup 1 --no-cpu --no-e2e"
python bench.py $SMALL > $O/${TAG}_bench_small.json 2> $O/${TAG}_bench_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv python bench.py $SMALL > $O/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scan_chunks|k_chunk_prefix|k_chunk_compact|k_parse|k_group|k_emit' -s 21 -c 8 -f -o $O/${TAG}_s2p python bench.py $SMALL > $O/${TAG}_ncu_full.log 2>&1
ls -la $O/${TAG}_*
