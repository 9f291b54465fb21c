#!/bin/bash
mkdir -p gpurun_out
python tools/host_overhead_probe.py > gpurun_out/r2c30_host.log 2>&1; echo "rc=$?"
echo done
