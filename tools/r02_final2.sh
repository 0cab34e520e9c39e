#!/bin/bash
# round-2 final measurement pass on one B200 (outputs gpurun_out/TAG_*; summaries are copied to profiles/ afterwards)
TAG=${1:-r02_final2}
O=gpurun_out
mkdir -p $O
(free -g; nproc; lscpu | grep -E "Model name|Socket|Core|Thread"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv) > $O/${TAG}_host.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 4 $O/${TAG}_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -n 3 $O/${TAG}_smoke.log
timeout 1500 python bench.py > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 600 $O/${TAG}_bench_c5.json
timeout 900 python bench.py --workload c3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; tail -c 300 $O/${TAG}_bench_c3.json
timeout 900 python bench.py --workload c4 --no-cpu-baseline > $O/${TAG}_bench_c4_n1.json 2> $O/${TAG}_bench_c4_n1.err; tail -c 300 $O/${TAG}_bench_c4_n1.json
timeout 600 python bench.py --workload c2 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; tail -c 300 $O/${TAG}_bench_c2.json
timeout 600 python bench.py --workload c2rock > $O/${TAG}_bench_c2rock.json 2> $O/${TAG}_bench_c2rock.err; tail -c 300 $O/${TAG}_bench_c2rock.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_ref.json 2>&1; tail -c 400 $O/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c5.csv python bench.py --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --workload c3 --steps 4 --warmup 20 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_c3_random.csv python bench.py --workload c3 --state random --steps 4 --warmup 6 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncu_c3_random.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 24 -c 2 -f -o $O/${TAG}_collide_c5 python bench.py --steps 2 --warmup 24 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull_c5.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 24 -c 2 -f -o $O/${TAG}_collide_c3 python bench.py --workload c3 --steps 2 --warmup 24 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull_c3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gradient_pack|k_chain_flat' -s 30 -c 6 -f -o $O/${TAG}_chain_c3_random python bench.py --workload c3 --state random --steps 2 --warmup 8 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull_chain.log 2>&1
ls -la $O | grep ${TAG} | wc -l
