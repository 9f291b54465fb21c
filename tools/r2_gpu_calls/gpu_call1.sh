#!/bin/bash
# first GPU call of round 2: parity, bench, tolerances, single-bag timeline, reduction counters
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
python tools/measure_tolerances.py > gpurun_out/r2c1_tol.log 2>&1; echo "tol rc=$?"
python tools/single_bag_probe.py 200 auto > gpurun_out/r2c1_single.log 2>&1
python tools/single_bag_probe.py 200 split >> gpurun_out/r2c1_single.log 2>&1
MCMIL_NO_PDL=1 python tools/single_bag_probe.py 200 auto >> gpurun_out/r2c1_single.log 2>&1
cat gpurun_out/r2c1_single.log
python tools/single_bag_probe.py 30 auto > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2c1_single_launches.csv python tools/single_bag_probe.py 30 auto > gpurun_out/r2c1_ncu1.log 2>&1
python tools/single_bag_probe.py 30 split > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2c1_single_launches_split.csv python tools/single_bag_probe.py 30 split > gpurun_out/r2c1_ncu2.log 2>&1
for wl in config2 config3 config4; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'softmax|welford|fused' -c 12 --csv --log-file gpurun_out/r2c1_red_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/r2c1_ncu_$wl.log 2>&1
done
echo done
