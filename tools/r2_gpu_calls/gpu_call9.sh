#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c9
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices or reduction_dispatch or config4 or config3" > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${P}_bench2.json 2> ${P}_bench2.err; echo "bench2 rc=$?"
tail -c 600 ${P}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > ${P}_ref2.json 2> ${P}_ref2.err; echo "ref2 rc=$?"
echo done
