set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/ab_all_tests.log 2>&1; tail -3 $O/ab_all_tests.log
python tools/shape_sweep.py 2>&1 | tail -22
python tools/bench_configs.py c4 2>&1 | tail -3
