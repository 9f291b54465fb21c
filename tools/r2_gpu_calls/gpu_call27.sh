#!/bin/bash
# compute-sanitizer memcheck over the small end-to-end cases (all row-kernel forms) + the stress script
mkdir -p gpurun_out
P=gpurun_out/r2c27
python tests/sanitize_case.py > ${P}_plain.log 2>&1; echo "plain rc=$?"; tail -1 ${P}_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_case.py > ${P}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 ${P}_memcheck.log
timeout 900 python tests/stress_gpu.py > ${P}_stress.log 2>&1; echo "stress rc=$?"; tail -2 ${P}_stress.log
echo done
