#!/bin/bash
# 2-GPU sanity of both bench arms as the driver launches them
mkdir -p gpurun_out
P=gpurun_out/r2c20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > ${P}_ref.json 2> ${P}_ref.err; echo "ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"; tail -c 300 ${P}_bench.err
echo done
