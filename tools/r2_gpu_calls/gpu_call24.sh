#!/bin/bash
# column kernel: double-buffered prefetch / occupancy / split variants at configs 2, 3, 4 (ncu durations)
mkdir -p gpurun_out
P=gpurun_out/r2c24
for v in base pf_u2 pf_u3 pf_u4 pf_u2_m5 u2_m6 ws3 pf_u2_ws2 pf_u4_ws2; do
  if [ $v = base ]; then unset MCMIL_LIB_PATH; else export MCMIL_LIB_PATH=build/variants/$v.so; fi
  for wl in config2 config3 config4; do
    case $v in ws3|pf_u2_ws2|pf_u4_ws2) [ $wl = config4 ] || continue;; esac
    CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
    $CMD > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford' -c 3 --csv --log-file ${P}_red_${v}_$wl.csv $CMD > ${P}_ncu_${v}_$wl.log 2>&1
  done
done
export MCMIL_LIB_PATH=build/variants/pf_u2.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dispatch or config4 or config3 or ragged or golden_inkernel" > ${P}_pytest_pf_u2.log 2>&1; echo "pytest pf_u2 rc=$?"; tail -2 ${P}_pytest_pf_u2.log
echo done
