#!/bin/bash
# One `gpurun --gpus N` call: multi-GPU parity tests (slabs vs single-domain oracle) and the weak-scaling bench line.
# usage: tools/mgpu.sh TAG N
TAG=${1:-r01_mgpu}
N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 4 $O/${TAG}_pytest.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 60 --warmup 10 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 60 --warmup 10 --no-cpu-baseline > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
  fi
  tail -c 2500 $O/${TAG}_bench_n$n.json; tail -n 5 $O/${TAG}_bench_n$n.err
done
ls -la $O
