#!/bin/bash
# hybrid chain (MFLBM_MARCH=2: fused kernel around the active tiles, flat list sweeps otherwise) against the list kernels, drainage state
TAG=${1:-r02_hybrid}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_march_gpu.py -q > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?"
tail -4 $O/${TAG}_pytest.log
run() {  # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 900 python bench.py --workload $wl --steps 40 --warmup 30 --no-cpu-baseline --no-active --no-e2e > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_${name}.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['roofline']['step_frac_of_roofline'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
}
run c3_lists c3 MFLBM_X=1
run c3_hybrid16 c3 MFLBM_MARCH=2
run c3_hybrid8 c3 MFLBM_MARCH=2 MFLBM_MARCH_LZR=8
run c3_hybrid32 c3 MFLBM_MARCH=2 MFLBM_MARCH_LZR=32
run c5_hybrid8 c5 MFLBM_MARCH=2 MFLBM_MARCH_LZR=8
run c5_hybrid16 c5 MFLBM_MARCH=2
