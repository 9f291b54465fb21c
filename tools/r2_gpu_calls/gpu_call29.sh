#!/bin/bash
# 8-GPU run of both bench arms as the driver launches them (scaling numbers of README / DESIGN)
mkdir -p gpurun_out
P=gpurun_out/r2c29
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > ${P}_bench_n$N.json 2> ${P}_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 300 ${P}_bench_n$N.err
echo done
