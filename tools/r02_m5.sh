#!/bin/bash
TAG=${1:-r02_m10}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 8 $O/${TAG}_pytest.log | cut -c1-300
timeout 900 python bench.py --workload c3 --no-cpu-baseline --no-active --no-e2e --steps 60 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 1000 $O/${TAG}_bench_c3.json
timeout 900 python bench.py --workload c5 --no-cpu-baseline --no-active --no-e2e --steps 60 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 1000 $O/${TAG}_bench_c5.json
