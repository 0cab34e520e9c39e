#!/bin/bash
# round 2, first measurement: GPU tests, c3 with the interface-rich legs, c2rock, launch lists + full capture in the random state
TAG=${1:-r02_m1}
O=gpurun_out
mkdir -p $O
(free -g; nproc; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv) > $O/${TAG}_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 3 $O/${TAG}_pytest.log
timeout 900 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 2500 $O/${TAG}_bench_c3.json
timeout 600 python bench.py --workload c2rock --no-cpu-baseline > $O/${TAG}_bench_c2rock.json 2> $O/${TAG}_bench_c2rock.err; tail -c 800 $O/${TAG}_bench_c2rock.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${TAG}_launches_c3_random.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c3_random.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 8 -c 2 -f -o $O/${TAG}_collide_c3_random python bench.py --workload c3 --state random --steps 2 --warmup 8 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull_c3_random.log 2>&1
ls -la $O | tail -12
