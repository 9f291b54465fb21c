#!/bin/bash
# fused single-launch reduction for short bags: parity, stress, single-bag latency, throughput mode
mkdir -p gpurun_out
P=gpurun_out/r2c36
timeout 900 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 ${P}_pytest.log
timeout 600 python tests/stress_gpu.py > ${P}_stress.log 2>&1; echo "stress rc=$?"; tail -2 ${P}_stress.log
python tests/sanitize_case.py > ${P}_san.log 2>&1; echo "san rc=$?"; tail -1 ${P}_san.log
python tools/single_bag_probe.py 300 x graph > ${P}_single.log 2>&1; cat ${P}_single.log
python tools/stream_probe.py 4,8 0 > ${P}_streams.log 2>&1; cat ${P}_streams.log
build/c_client > ${P}_c.log 2>&1; echo "c_client rc=$?"; tail -2 ${P}_c.log
echo done
