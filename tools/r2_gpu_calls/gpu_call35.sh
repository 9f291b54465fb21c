#!/bin/bash
# flakiness check: the GPU suite four times, the stress script twice
mkdir -p gpurun_out
P=gpurun_out/r2c35
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_$i.log 2>&1; echo "pytest $i rc=$?"; tail -1 ${P}_pytest_$i.log
done
for i in 1 2; do timeout 600 python tests/stress_gpu.py > ${P}_stress_$i.log 2>&1; echo "stress $i rc=$?"; tail -1 ${P}_stress_$i.log; done
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
echo done
