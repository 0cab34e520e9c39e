#!/bin/bash
# last check of the round on the final tree: full GPU test suite, smoke, the default bench line exactly as the driver runs it
TAG=${1:-r02_last}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 4 $O/${TAG}_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -n 3 $O/${TAG}_smoke.log
timeout 1200 python bench.py > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 700 $O/${TAG}_bench_c5.json
