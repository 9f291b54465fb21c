#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c28
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
build/c_client > ${P}_c_client.log 2>&1; echo "c_client rc=$?"; cat ${P}_c_client.log
echo done
