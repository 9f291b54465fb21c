#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c26
python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"; tail -c 400 ${P}_bench.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --e2e-streams 4 > ${P}_bench_s4.json 2> ${P}_bench_s4.err; echo "bench rc=$?"
echo done
