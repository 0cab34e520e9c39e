#!/bin/bash
# One gpurun call: device geometry preprocessing parity + the configs[4] slab (1536x1536x192) on one GPU.
TAG=${1:-r01_v9}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_geometry_gpu.py -q > $O/${TAG}_pytest_geo.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_geo.log
tail -n 30 $O/${TAG}_pytest_geo.log
timeout 1100 python bench.py --workload c5 --steps 20 --warmup 6 --no-cpu-baseline > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
tail -c 1800 $O/${TAG}_bench_c5.json; tail -n 5 $O/${TAG}_bench_c5.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
