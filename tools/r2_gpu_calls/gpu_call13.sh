#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c13
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${P}_ref.json 2> ${P}_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"; tail -c 300 ${P}_bench.err
echo done
