set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "search or generate" > $O/ab_all_tests.log 2>&1; tail -2 $O/ab_all_tests.log
python tools/shape_sweep.py 2>&1 | tail -22
