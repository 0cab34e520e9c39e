#!/bin/bash
TAG=${1:-r02_m12}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_yperiodic_gpu.py -q -m gpu > $O/${TAG}_pytest_y.log 2>&1; tail -n 30 $O/${TAG}_pytest_y.log | cut -c1-600
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_inputs_gpu.py tests/test_checkpoint_gpu.py -q -m gpu > $O/${TAG}_pytest_b.log 2>&1; tail -n 4 $O/${TAG}_pytest_b.log | cut -c1-300
