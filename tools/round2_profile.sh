#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun from the repo root; outputs in gpurun_out/, summaries
# are made from them with tools/summarize_ncu.py and committed under profiles/).
#   1. bench.py plain, both arms            2. ncu launch list of the same bench command
#   3. ncu --set full of the screening kernel, the fp64 headline kernel, K2 on supplied designs (both kernels)
#   4. K2 rates on supplied designs (tools/score_micro.py)
set -u
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
cat $O/${TAG}_bench_n1.json | cut -c1-600
python tools/score_micro.py > $O/${TAG}_score_micro.txt 2>&1
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --candidates 16777216 --no-cpu-baseline --no-configs"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_bench_launches.csv $B > $O/${TAG}_ncu_launch.log 2>&1
$NCU --set full --import-source on -k regex:search_screen -s 3 -c 1 -o $O/${TAG}_search_screen $B > $O/${TAG}_ncu_screen.log 2>&1
$NCU --set full --import-source on -k regex:search_cyclic -s 7 -c 1 -o $O/${TAG}_search_cyclic $B > $O/${TAG}_ncu_cyclic.log 2>&1
$NCU --set full --import-source on -k regex:search_cyclic -c 1 -o $O/${TAG}_score_supplied_cyclic python tools/score_micro.py 50 3 > $O/${TAG}_ncu_k2a.log 2>&1
$NCU --set full --import-source on -k regex:pair_ -c 1 -o $O/${TAG}_score_supplied_generic python tools/score_micro.py 500 10 > $O/${TAG}_ncu_k2b.log 2>&1
ls -la $O | tail -12
