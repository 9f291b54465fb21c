#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c4
for path in auto split; do
  for skip in none reduce proj; do
    if [ $skip = none ]; then python tools/single_bag_probe.py 300 $path graph; else MCMIL_EXP_SKIP=$skip python tools/single_bag_probe.py 300 $path graph; fi 2>&1 | sed "s/^/skip=$skip /"
  done
done > ${P}_single.log 2>&1
cat ${P}_single.log
python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
for wl in config2 config3 config4; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'softmax|welford|fused' -c 12 --csv --log-file ${P}_red_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_$wl.log 2>&1
done
echo done
