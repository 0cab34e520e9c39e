#!/bin/bash
TAG=${1:-r02_k4c}
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c5.csv python bench.py --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c5.log 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py tests/test_reference_inputs_gpu.py tests/test_yperiodic_gpu.py -q -m gpu > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
