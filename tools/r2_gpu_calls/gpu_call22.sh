#!/bin/bash
# row kernel family (CTA per row, streaming warps) and sample-split variants of the column kernel
mkdir -p gpurun_out
P=gpurun_out/r2c22
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
for v in base prev_rows ws2 ws8 ws9 ws16; do
  if [ $v = base ]; then unset MCMIL_LIB_PATH; else export MCMIL_LIB_PATH=build/variants/$v.so; fi
  for wl in config4; do
    CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
    $CMD > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford|softmax_rows' -c 6 --csv --log-file ${P}_red_${v}_$wl.csv $CMD > ${P}_ncu_${v}_$wl.log 2>&1
  done
  python tools/single_bag_probe.py 300 x graph > ${P}_single_${v}.log 2>&1
done
unset MCMIL_LIB_PATH
for wl in config2 config3; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
  $CMD > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford|softmax_rows' -c 6 --csv --log-file ${P}_red_base_$wl.csv $CMD > ${P}_ncu_base_$wl.log 2>&1
done
echo done
