#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c11
python tools/stream_probe.py > ${P}_streams.log 2>&1; cat ${P}_streams.log
# profiles: launch list of the bench command, full counters of the three hot kernels, DRAM traffic of the projection
CMD="python bench.py --steps 2 --warmup 1 --bags-per-step 32 --no-e2e --no-cpu --no-extras"
$CMD > ${P}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file ${P}_launches.csv $CMD > ${P}_ncu_l.log 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'proj_tc|softmax_rows|welford_cols' -s 6 -c 3 -o ${P}_full $CMD > ${P}_ncu_f.log 2>&1
CMD2="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
$CMD2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'proj_tc' -s 3 -c 2 --csv --log-file ${P}_traffic.csv $CMD2 > ${P}_ncu_t.log 2>&1
for wl in config2 config4; do
  python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'softmax|welford' -c 8 --csv --log-file ${P}_red_$wl.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > ${P}_ncu_$wl.log 2>&1
done
ls -la gpurun_out/r2c11*
echo done
