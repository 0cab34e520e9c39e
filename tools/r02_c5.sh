#!/bin/bash
# c5 slab (1536x1536x192) on one GPU: bench line with the interface-rich legs + launch list of the drainage state
TAG=${1:-r02_c5}
O=gpurun_out
mkdir -p $O
timeout 1500 python bench.py --workload c5 --no-cpu-baseline > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 3000 $O/${TAG}_bench_c5.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c5.csv python bench.py --workload c5 --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c5.log 2>&1
ls -la $O | tail -5
