#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c8
for v in base nopad tpc2u2 tpc4u2 tpc4u4 tpc1u4 tpc8u2; do
for wl in config2 config3 config4; do
  if [ $v = base ]; then unset MCMIL_LIB_PATH; else export MCMIL_LIB_PATH=build/variants/$v.so; fi
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'welford' -c 4 --csv --log-file ${P}_red_${v}_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_${v}_$wl.log 2>&1
done
done
unset MCMIL_LIB_PATH
python bench.py --workload config2 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'welford' -s 2 -c 1 -o ${P}_welford python bench.py --workload config2 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_full.log 2>&1
python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
echo done
