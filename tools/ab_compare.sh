#!/bin/bash
# One gpurun call: GPU parity tests, then an A/B comparison of one environment knob on the C3 and the 1536x1536x192 slab
# workloads (used for r01_v10: MFLBM_STATIC_BC_TILES=1, r01_v11: K4 list vs shared memory).
# usage: tools/ab_compare.sh TAG KNOB=VALUE
TAG=${1:-r01_ab}
KNOB=${2:-MFLBM_K4_SMEM=1}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 12 $O/${TAG}_pytest.log
for wl in c3 c5; do
  timeout 900 python bench.py --workload $wl --steps 40 --warmup 10 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  env $KNOB timeout 900 python bench.py --workload $wl --steps 40 --warmup 10 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_${wl}_knob.json 2> $O/${TAG}_bench_${wl}_knob.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_bench_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]
            print(f, "MLUPS %.0f ms/step %.3f kernel %.3f quiet %.3f step_frac %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_step"], d["config"]["quiet_tile_fraction"] or 0, r["step_frac_of_roofline"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_c5.csv python bench.py --workload c5 --steps 4 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c5.log 2>&1
ls $O | grep $TAG
