#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c23
timeout 300 build/colread_bench 1000 2 16384 > ${P}_colread_config4.log 2>&1; echo "rc=$?"
timeout 300 build/colread_bench 100 2 262144 > ${P}_colread_config2.log 2>&1; echo "rc=$?"
echo done
