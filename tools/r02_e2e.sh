#!/bin/bash
# K4+K5 fusion on the tile-driven path + streamed e2e: tests, then c3 and c5 (drainage) with the e2e leg
TAG=${1:-r02_e2e}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_output_gpu.py tests/test_march_gpu.py tests/test_parity_gpu.py tests/test_spherepack_gpu.py tests/test_reference_inputs_gpu.py tests/test_checkpoint_gpu.py tests/test_golden_ref.py -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?"
tail -6 $O/${TAG}_pytest.log
timeout 600 python bench.py --workload c3 --steps 40 --warmup 30 --no-cpu-baseline --no-active > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_c3.json')); print('c3', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_roofline'], d['e2e'])"
timeout 600 python bench.py --workload c3 --steps 40 --warmup 30 --no-cpu-baseline --no-active --e2e-blocking > $O/${TAG}_bench_c3_blocking.json 2> $O/${TAG}_bench_c3_blocking.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_c3_blocking.json')); print('c3 blocking', d['value'], d['ms_per_step'], d['e2e'])"
timeout 900 python bench.py --steps 30 --warmup 20 --no-cpu-baseline --no-active > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_c5.json')); print('c5', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_roofline'], d['e2e'], d['clocks'])"
timeout 300 python bench.py --workload c2 --steps 200 --warmup 50 --no-cpu-baseline --no-active > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_roofline'], d['e2e'])"
