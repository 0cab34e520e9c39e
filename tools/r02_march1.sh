#!/bin/bash
# first GPU contact of the march kernel: its own tests, the sparse parity suites, launch lists of C3 (random phi / drainage)
TAG=${1:-r02_march1}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_march_gpu.py -x -q > $O/${TAG}_pytest_march.log 2>&1; echo "march tests rc=$?" | tee -a $O/${TAG}_summary.txt
tail -5 $O/${TAG}_pytest_march.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py tests/test_reference_inputs_gpu.py tests/test_yperiodic_gpu.py -x -q > $O/${TAG}_pytest_parity.log 2>&1; echo "parity tests rc=$?" | tee -a $O/${TAG}_summary.txt
tail -5 $O/${TAG}_pytest_parity.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c3_random.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_random.log 2>&1
timeout 600 python bench.py --workload c3 --steps 20 --warmup 30 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
tail -c 1500 $O/${TAG}_bench_c3.json
