#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun from the repo root):
#   1. bench.py, plain (the numbers that count), both arms
#   2. ncu launch list of the same bench command (kernel shares of the step)
#   3. ncu --set full captures of the dominant kernel of each path (search, CTA search, anneal, order)
#   4. the BASELINE.json configs C1-C5
# Outputs land in gpurun_out/; summaries are made from them with tools/summarize_ncu.py.
set -u
TAG=${1:-r1f}
O=gpurun_out
mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
cat $O/${TAG}_bench_n1.json
python tools/bench_configs.py > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err
cat $O/${TAG}_configs.jsonl
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --candidates 16777216 --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
$NCU --set full --import-source on -k regex:search_cyclic -s 3 -c 1 -o $O/${TAG}_search_cyclic \
    python bench.py --steps 2 --warmup 3 --candidates 16777216 --no-cpu-baseline > $O/${TAG}_ncu_search.log 2>&1
$NCU --set full --import-source on -k regex:search_cta -c 1 -o $O/${TAG}_search_cta \
    python tools/bench_configs.py c3 > $O/${TAG}_ncu_cta.log 2>&1
$NCU --set full --import-source on -k regex:anneal_swap_pipe -s 2 -c 1 -o $O/${TAG}_anneal_pipe \
    python tools/anneal_prof.py 2000 > $O/${TAG}_ncu_anneal.log 2>&1
$NCU --set full --import-source on -k regex:order_cluster -s 1 -c 1 -o $O/${TAG}_order_cluster \
    python tools/bench_configs.py c5 > $O/${TAG}_ncu_order.log 2>&1
$NCU --set full --import-source on -k regex:anneal_maximin_swap -c 1 -o $O/${TAG}_anneal_maximin \
    python tools/anneal_maximin_prof.py 500 > $O/${TAG}_ncu_anneal_maximin.log 2>&1
$NCU --set full --import-source on -k regex:search_warp -s 1 -c 1 -o $O/${TAG}_search_warp \
    python tools/search_prof.py 100 10 1000000 > $O/${TAG}_ncu_warp.log 2>&1
$NCU --set full --import-source on -k regex:search_small -s 1 -c 1 -o $O/${TAG}_search_small \
    python tools/search_prof.py 50 8 8000000 > $O/${TAG}_ncu_small.log 2>&1
python tools/order_ab.py > $O/${TAG}_order_ab.txt 2>&1
python tools/shape_sweep.py > $O/${TAG}_shape_sweep_maxpro.txt 2>&1
python tools/shape_sweep.py maximin > $O/${TAG}_shape_sweep_maximin.txt 2>&1
ls -la $O | tail -20
