#!/bin/bash
# one multi-GPU bench line: tools/r02_scale.sh TAG N WORKLOAD [extra bench.py flags]
TAG=$1; N=$2; W=$3; shift 3
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo_n${N}.txt 2>&1
if [ "$N" = "1" ]; then
  timeout 1700 python bench.py --gpus 1 --workload $W "$@" > $O/${TAG}_bench_${W}_n${N}.json 2> $O/${TAG}_bench_${W}_n${N}.err
else
  timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $W "$@" > $O/${TAG}_bench_${W}_n${N}.json 2> $O/${TAG}_bench_${W}_n${N}.err
fi
echo "exit $?"; tail -c 3500 $O/${TAG}_bench_${W}_n${N}.json; tail -n 5 $O/${TAG}_bench_${W}_n${N}.err
