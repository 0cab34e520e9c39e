#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (c3, c2, reference arm), ncu launch lists and full captures of k_collide.
# usage: tools/gpu_measure.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-r01_v13}
O=gpurun_out
mkdir -p $O
(free -g; nproc; lscpu | grep -E "Model name|Socket|Core|Thread"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv) > $O/${TAG}_host.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 3 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 1500 $O/${TAG}_bench_c3.json
timeout 600 python bench.py --workload c2 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; tail -c 1500 $O/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_ref.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 4 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --workload c2 --steps 4 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_c2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 6 -c 2 -f -o $O/${TAG}_collide_c3 python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncufull_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 6 -c 2 -f -o $O/${TAG}_collide_c2 python bench.py --workload c2 --steps 2 --warmup 6 --no-cpu-baseline --no-e2e > $O/${TAG}_ncufull_c2.log 2>&1
ls -la $O
