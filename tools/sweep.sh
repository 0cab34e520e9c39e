#!/bin/bash
# One gpurun call: parity tests, then a sweep of kernel tunings (library variants + environment knobs), then ncu captures.
# usage: tools/sweep.sh TAG
TAG=${1:-r01_v4}
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
tail -5 $O/${TAG}_pytest.log
MFLBM_PIPE=2 MFLBM_PIPE_GRID=24 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py -q > $O/${TAG}_pytest_pipe.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_pipe.log
tail -5 $O/${TAG}_pytest_pipe.log
run base c3 A=1
run base c2 A=1
run pipe1 c3 MFLBM_PIPE=1
run pipe2 c3 MFLBM_PIPE=2
run pipe1 c2 MFLBM_PIPE=1
run pipe2 c2 MFLBM_PIPE=2
run pipe1_pf0 c2 MFLBM_PIPE=1 MFLBM_PF_DIST=0
run skew4352 c3 MFLBM_ARRAY_SKEW=4352
run skew69888 c3 MFLBM_ARRAY_SKEW=69888
run skew4352_pipe1 c3 MFLBM_ARRAY_SKEW=4352 MFLBM_PIPE=1
run skew4352 c2 MFLBM_ARRAY_SKEW=4352
cat $L
counters base c3 A=1
counters skew c3 MFLBM_ARRAY_SKEW=4352
counters pipe1 c3 MFLBM_PIPE=1
counters base c2 A=1
counters pipe1 c2 MFLBM_PIPE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide_pipe -s 2 -c 1 -f -o $O/${TAG}_pipe_c3 env MFLBM_PIPE=1 python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-e2e > $O/${TAG}_ncufull_pipe_c3.log 2>&1
ls -la $O
