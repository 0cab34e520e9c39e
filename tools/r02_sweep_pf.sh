#!/bin/bash
# odd-step metadata prefetch (MFLBM_PF_MODE=2) distance sweep on c3; one JSON summary line per setting
TAG=${1:-r02_pf}
O=gpurun_out
mkdir -p $O
run() { # label, env...
  L=$1; shift
  env "$@" timeout 600 python bench.py --workload c3 --steps 60 --warmup 10 --no-active --no-e2e --no-cpu-baseline 2>$O/${TAG}_$L.err | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['step_frac_of_roofline'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" | tee -a $O/${TAG}_sweep.txt
}
run base MFLBM_PF_DIST=0
run m2_8k MFLBM_PF_MODE=2 MFLBM_PF_DIST=8192
run m2_32k MFLBM_PF_MODE=2 MFLBM_PF_DIST=32768
run m2_128k MFLBM_PF_MODE=2 MFLBM_PF_DIST=131072
run m2_512k MFLBM_PF_MODE=2 MFLBM_PF_DIST=524288
run base2 MFLBM_PF_DIST=0
