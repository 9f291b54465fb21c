#!/bin/bash
# throughput mode: does the overlap depend on the number of hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS)?
mkdir -p gpurun_out
P=gpurun_out/r2c32
sed -i 's/itertools.product((4, 6, 8), (0, 8, 16, 28))/itertools.product((4, 8), (0,))/' tools/stream_probe.py
for mc in default 32; do
  if [ $mc = default ]; then unset CUDA_DEVICE_MAX_CONNECTIONS; else export CUDA_DEVICE_MAX_CONNECTIONS=$mc; fi
  python tools/stream_probe.py > ${P}_probe_$mc.log 2>&1; echo "probe $mc rc=$?"; cat ${P}_probe_$mc.log
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs > ${P}_bench_$mc.json 2> ${P}_bench_$mc.err; echo "bench $mc rc=$?"
done
echo done
