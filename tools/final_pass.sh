#!/bin/bash
# What the round ends with on one B200 (run under gpurun from the repo root): smoke(), the GPU test suite,
# then tools/round_profile.sh (bench, configs, ncu captures, sweeps).
TAG=${1:-r1j}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/${TAG}_pytest_gpu.txt
bash tools/round_profile.sh $TAG > $O/${TAG}_round.log 2>&1
tail -3 $O/${TAG}_round.log
cat $O/${TAG}_order_ab.txt
