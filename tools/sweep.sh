#!/bin/bash
# One gpurun call: parity tests, then a sweep of kernel tunings (library variants + environment knobs), then ncu captures.
# usage: tools/sweep.sh TAG
TAG=${1:-r01_v6}
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
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 5 $O/${TAG}_pytest.log
run base c3 A=1
run static_bc_tiles c3 MFLBM_STATIC_BC_TILES=1
run pf2 c3 MFLBM_PF_DIST=65536 MFLBM_PF_MODE=2
run pf2_32k c3 MFLBM_PF_DIST=32768 MFLBM_PF_MODE=2
run blk64 c3 MFLBM_LIB_VARIANT=blk64
run blk32 c3 MFLBM_LIB_VARIANT=blk32
run blk64_pf2 c3 MFLBM_LIB_VARIANT=blk64 MFLBM_PF_DIST=65536 MFLBM_PF_MODE=2
run base c2 A=1
run pf2 c2 MFLBM_PF_MODE=2
run blk64 c2 MFLBM_LIB_VARIANT=blk64
run blk32 c2 MFLBM_LIB_VARIANT=blk32
cat $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 4 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c3.log 2>&1
ls -la $O
