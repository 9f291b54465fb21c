#!/bin/bash
# throughput mode: does the overlap depend on the number of hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS)?
mkdir -p gpurun_out
P=gpurun_out/r2c32
for mc in default 32; do
  if [ $mc = default ]; then unset CUDA_DEVICE_MAX_CONNECTIONS; else export CUDA_DEVICE_MAX_CONNECTIONS=$mc; fi
  python tools/stream_probe.py 4,8 0 > ${P}_probe_$mc.log 2>&1; echo "probe $mc rc=$?"; cat ${P}_probe_$mc.log
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > ${P}_bench_$mc.json 2> ${P}_bench_$mc.err; echo "bench $mc rc=$?"
done
echo done
