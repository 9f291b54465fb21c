#!/bin/bash
mkdir -p gpurun_out
python tools/stream_probe.py > gpurun_out/r2c14_streams.log 2>&1; cat gpurun_out/r2c14_streams.log
echo done
