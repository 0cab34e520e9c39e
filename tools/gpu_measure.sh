#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (c3, c2, reference arm), ncu launch lists and full captures of k_collide.
# usage: tools/gpu_measure.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-r01_v2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; tail -c 1500 gpurun_out/${TAG}_bench_c3.json
timeout 600 python bench.py --workload c2 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -c 1500 gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --workload c2 --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 4 -c 2 -f -o gpurun_out/${TAG}_collide_c3 python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_ncufull_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 4 -c 2 -f -o gpurun_out/${TAG}_collide_c2 python bench.py --workload c2 --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_ncufull_c2.log 2>&1
ls -la gpurun_out
