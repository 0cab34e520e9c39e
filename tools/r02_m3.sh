#!/bin/bash
TAG=${1:-r02_m8}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 12 $O/${TAG}_pytest.log
(timeout 600 python tools/debug/case5.py; MFLBM_NO_TILES=1 timeout 600 python tools/debug/case5.py) > $O/${TAG}_case5.log 2>&1; cat $O/${TAG}_case5.log | cut -c1-400
timeout 900 python bench.py --workload c3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 900 $O/${TAG}_bench_c3.json
MFLBM_NO_SPEC=1 timeout 900 python bench.py --workload c3 --no-cpu-baseline --no-active --no-e2e > $O/${TAG}_bench_c3_nospec.json 2> $O/${TAG}_bench_c3_nospec.err; tail -c 600 $O/${TAG}_bench_c3_nospec.json
timeout 900 python bench.py --workload c5 --no-cpu-baseline --no-active --no-e2e > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 900 $O/${TAG}_bench_c5.json
MFLBM_NO_SPEC=1 timeout 900 python bench.py --workload c5 --no-cpu-baseline --no-active --no-e2e > $O/${TAG}_bench_c5_nospec.json 2> $O/${TAG}_bench_c5_nospec.err; tail -c 600 $O/${TAG}_bench_c5_nospec.json
