#!/bin/bash
# programmatic dependent launch of the collision / wrap kernels (MFLBM_PDL=1): parity tests with it, singlephase C2 A/B, C3 check
TAG=${1:-r02_pdl}
O=gpurun_out
mkdir -p $O
MFLBM_PDL=1 timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py -q -x > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?"
tail -2 $O/${TAG}_pytest.log
run() {
  name=$1; wl=$2; st=$3; shift 3
  env "$@" timeout 200 python bench.py --workload $wl --steps $st --warmup 50 --no-cpu-baseline --no-active > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
  python -c "
import json; d=json.load(open('$O/${TAG}_bench_${name}.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['step_frac_of_roofline'],4), 'e2e', round(d['e2e']['value'],1))"
}
run c2_off c2 400 MFLBM_PDL=0
run c2_pdl c2 400 MFLBM_PDL=1
run c3_pdl c3 40 MFLBM_PDL=1
