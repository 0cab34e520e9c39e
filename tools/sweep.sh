#!/bin/bash
# One gpurun call: a sweep of kernel tunings (environment knobs and `make variant` library builds), one short bench line
# each; the sweeps behind profiles/r01_v3..v6_sweep.txt were produced with earlier lists of this script.
# usage: tools/sweep.sh TAG            (edit the list at the bottom)
TAG=${1:-r01_sweep}
O=gpurun_out
mkdir -p $O
L=$O/${TAG}_sweep.log
: > $L
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  local steps=40; [ "$wl" = c2 ] && steps=200
  echo "== $label $wl $*" >> $L
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --warmup 10 --no-cpu-baseline --no-e2e 2>> $O/${TAG}_sweep.err \
    | python -c 'import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d["roofline"]; print("   MLUPS %.0f  ms/step %.4f  k_collide ms/step %.4f  kernel frac %.3f  step frac %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_step"], r["frac"], r["step_frac_of_roofline"]))' >> $L
}
counters() {  # label, workload, env...: DRAM / L2 counters of one odd + one even collision launch
  local label=$1 wl=$2; shift 2
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_op_read_hit_rate.pct,lts__t_sector_op_write_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
    --clock-control none -k regex:k_collide -s 4 -c 2 --csv --log-file $O/${TAG}_ctr_${label}_${wl}.csv python bench.py --workload $wl --steps 2 --warmup 4 --no-cpu-baseline --no-e2e > $O/${TAG}_ctr_${label}_${wl}.log 2>&1
}
run base c3 A=1
run base c2 A=1
run pf64k c3 MFLBM_PF_DIST=65536
run pf0 c2 MFLBM_PF_DIST=0
run no_tiles c3 MFLBM_NO_TILES=1
run static_ghost_tiles c3 MFLBM_STATIC_BC_TILES=1
run k4_smem c3 MFLBM_K4_SMEM=1
cat $L
counters base c3 A=1
counters base c2 A=1
ls -la $O | grep $TAG
