#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c10
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"; tail -c 400 ${P}_bench.err
for wl in config2 config3 config4; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'softmax|welford' -c 8 --csv --log-file ${P}_red_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_$wl.log 2>&1
done
echo done
