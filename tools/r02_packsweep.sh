#!/bin/bash
TAG=${1:-r02_pack}
O=gpurun_out
mkdir -p $O
for v in "" pack2 pack4 pack5; do
MFLBM_LIB_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${TAG}_launches_${v:-base}.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_${v:-base}.log 2>&1
done
