#!/bin/bash
# flat sweeps of the list kernels: K4+K5 fusion and brick orders (C3, random phi), launch lists per variant; tests first
TAG=${1:-r02_brick}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_yperiodic_gpu.py -q > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?"
tail -6 $O/${TAG}_pytest.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_${name}.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_${name}.log 2>&1
  python tools/ncu_summary.py launches $O/${TAG}_launches_${name}.csv | grep -E "k_chain_flat|k_gradient_pack_all|k_collide" | head -8
}
echo "== nok45 raster"; run nok45 MFLBM_NO_K45=1
echo "== raster"; run raster MFLBM_X=1
echo "== 32,8,4"; run b32x8x4 MFLBM_BRICK=32,8,4
echo "== 64,4,4"; run b64x4x4 MFLBM_BRICK=64,4,4
echo "== 0,4,4"; run rows4x4 MFLBM_BRICK=0,4,4
