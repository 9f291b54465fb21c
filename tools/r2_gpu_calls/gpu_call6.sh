#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c6
python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras"
for v in base hset2 base hset2; do
  if [ $v = base ]; then $B; else MCMIL_LIB_PATH=build/variants/$v.so $B; fi 2>${P}_ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); r=b['roofline']
print('$v', 'kernel_ms %.3f frac_burst %.3f value %.0f clocks %s'%(r['kernel_ms'], r['frac_of_burst'], b['value'], b['clocks']['sm_mhz']))"
done > ${P}_ab.log 2>&1
cat ${P}_ab.log
for v in base tpc1u4 tpc1u8 tpc2u8 tpc4u4 tpc4u8; do
for wl in config2 config3 config4; do
  if [ $v = base ]; then unset MCMIL_LIB_PATH; else export MCMIL_LIB_PATH=build/variants/$v.so; fi
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford' -c 4 --csv --log-file ${P}_red_${v}_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_${v}_$wl.log 2>&1
done
done
unset MCMIL_LIB_PATH
python tools/single_bag_probe.py 300 split graph > ${P}_single.log 2>&1; cat ${P}_single.log
echo done
