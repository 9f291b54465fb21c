#!/bin/bash
# final reduction captures (row kernel family, split rule) + tests + bench
mkdir -p gpurun_out
P=gpurun_out/r2c25
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 ${P}_pytest.log
for wl in config2 config3 config4; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
  $CMD > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford|softmax_rows' -c 6 --csv --log-file ${P}_red_base_$wl.csv $CMD > ${P}_ncu_base_$wl.log 2>&1
done
python tools/single_bag_probe.py 300 x graph > ${P}_single.log 2>&1; tail -2 ${P}_single.log
python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
echo done
