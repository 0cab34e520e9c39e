#!/bin/bash
TAG=${1:-r02_k4}
O=gpurun_out
mkdir -p $O
for m in 0 2; do
MFLBM_K4_SMEM=$m timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c5_k4m$m.csv python bench.py --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c5_k4m$m.log 2>&1
done
MFLBM_K4_SMEM=2 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py -q -m gpu > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
