#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c19
timeout 900 python tests/stress_gpu.py 7 40 > ${P}_stress.log 2>&1; echo "stress rc=$?"; tail -12 ${P}_stress.log
echo done
