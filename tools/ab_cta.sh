set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search" > $O/ab_cta_tests.log 2>&1; tail -3 $O/ab_cta_tests.log
for v in 0 1; do echo "MAXPRO_CTA_PIPE=$v"; MAXPRO_CTA_PIPE=$v python tools/bench_configs.py c3 2>&1 | tail -1; done
python tools/shape_sweep.py 2>&1 | tail -8
python tools/shape_sweep.py maximin 2>&1 | tail -8
