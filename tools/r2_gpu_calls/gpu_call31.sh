#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c31
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
python tools/host_overhead_probe.py > ${P}_host.log 2>&1; echo "rc=$?"; head -3 ${P}_host.log
echo done
