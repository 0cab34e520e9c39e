#!/bin/bash
# GPU tests (all, no -x), c3 bench with interface-rich legs, launch lists of both states
TAG=${1:-r02_m7}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 4 $O/${TAG}_pytest.log
timeout 900 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 1200 $O/${TAG}_bench_c3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --workload c3 --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c3_random.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c3_random.log 2>&1
