#!/bin/bash
# K7 + packing alone in brick order (cell and node index as independent arrays), C3 random phi, launch lists
TAG=${1:-r02_brick7}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_march_gpu.py -q -k flat_sweeps > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?"
tail -3 $O/${TAG}_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_${name}.csv python bench.py --workload c3 --state random --steps 3 --warmup 5 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_${name}.log 2>&1
  python tools/ncu_summary.py launches $O/${TAG}_launches_${name}.csv | grep -E "k_gradient_pack_all" | head -2
}
echo "== raster"; run raster MFLBM_X=1
echo "== 96,4,2"; run k96x4x2 MFLBM_BRICK7=96,4,2
echo "== 128,2,2"; run k128x2x2 MFLBM_BRICK7=128,2,2
echo "== 64,4,4"; run k64x4x4 MFLBM_BRICK7=64,4,4
