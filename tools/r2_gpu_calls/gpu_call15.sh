#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c15
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > ${P}_bench8.json 2> ${P}_bench8.err; echo "bench8 rc=$?"
tail -c 500 ${P}_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > ${P}_bench4.json 2> ${P}_bench4.err; echo "bench4 rc=$?"
echo done
