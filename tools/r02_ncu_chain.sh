#!/bin/bash
# full ncu capture of the gradient-chain kernels in the interface-rich state (c3, seeded option-6 random phi)
TAG=${1:-r02_chain}
O=gpurun_out
mkdir -p $O
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_gradient_pack|k_chain_flat' -s 30 -c 5 -f -o $O/${TAG}_chain_c3_random python bench.py --workload c3 --state random --steps 2 --warmup 8 --no-cpu-baseline --no-e2e --no-active > $O/${TAG}_ncufull_chain.log 2>&1
ls -la $O/${TAG}_chain_c3_random.ncu-rep
