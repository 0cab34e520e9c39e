#!/bin/bash
# march kernel: tests after the dead-storage mask fix, full ncu capture of one flat launch (C3, random phi)
TAG=${1:-r02_march2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_march_gpu.py -x -q > $O/${TAG}_pytest_march.log 2>&1; echo "march tests rc=$?"
tail -3 $O/${TAG}_pytest_march.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_spherepack_gpu.py tests/test_reference_inputs_gpu.py tests/test_yperiodic_gpu.py tests/test_checkpoint_gpu.py tests/test_output_gpu.py -q > $O/${TAG}_pytest_parity.log 2>&1; echo "parity tests rc=$?"
tail -8 $O/${TAG}_pytest_parity.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march' -s 10 -c 2 -f -o $O/${TAG}_march_c3_random python bench.py --workload c3 --state random --steps 2 --warmup 8 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull.log 2>&1
ls -la $O/${TAG}_march_c3_random.ncu-rep
