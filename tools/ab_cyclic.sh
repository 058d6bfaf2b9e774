set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/ab_tests.log 2>&1; tail -3 $O/ab_tests.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ab_bench.json 2> $O/ab_bench.err; python -c "
import json;d=json.load(open('$O/ab_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['clocks'],d.get('fp32_screen',{}).get('value'))"
python tools/bench_configs.py > $O/ab_configs.jsonl 2> $O/ab_configs.err; cat $O/ab_configs.jsonl
