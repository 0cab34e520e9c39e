#!/bin/bash
# One `gpurun --gpus N` call: multi-GPU parity tests (slabs vs single-domain oracle) and the weak-scaling bench line at N.
# usage: tools/mgpu.sh TAG N [workloads...]
TAG=${1:-r01_mgpu}
N=${2:-2}
shift 2
WLS=${@:-c3}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -n 4 $O/${TAG}_pytest.log
for wl in $WLS; do
  NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $wl --steps 60 --warmup 10 --no-cpu-baseline > $O/${TAG}_bench_${wl}_n$N.json 2> $O/${TAG}_bench_${wl}_n$N.err
  tail -c 2500 $O/${TAG}_bench_${wl}_n$N.json; tail -n 5 $O/${TAG}_bench_${wl}_n$N.err
done
ls -la $O | grep $TAG
