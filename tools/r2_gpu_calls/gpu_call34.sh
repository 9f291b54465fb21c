#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c34
python tools/stream_probe.py 2,4,8 0 > ${P}_probe.log 2>&1; echo "probe rc=$?"; cat ${P}_probe.log
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 ${P}_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
echo done
