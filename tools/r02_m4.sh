#!/bin/bash
TAG=${1:-r02_m9}
O=gpurun_out
mkdir -p $O
(cd _old && timeout 600 python tools/debug/case5.py) > $O/${TAG}_case5_old.log 2>&1; cut -c1-300 $O/${TAG}_case5_old.log
timeout 900 python bench.py --workload c3 --no-cpu-baseline --no-active --no-e2e --steps 60 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 1000 $O/${TAG}_bench_c3.json
timeout 900 python bench.py --workload c5 --no-cpu-baseline --no-active --no-e2e --steps 60 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 1000 $O/${TAG}_bench_c5.json
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "speculative or quiet" > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
