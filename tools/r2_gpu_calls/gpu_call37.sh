#!/bin/bash
# final profiles of the shipped binary: launch list + full counters of the three hot kernels (same command as r2c11)
mkdir -p gpurun_out
P=gpurun_out/r2c37
CMD="python bench.py --steps 2 --warmup 1 --bags-per-step 32 --no-e2e --no-cpu --no-extras"
$CMD > ${P}_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file ${P}_launches.csv $CMD > ${P}_ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'proj_tc|softmax_rows|welford_cols' -s 6 -c 3 -o ${P}_full $CMD > ${P}_ncu_f.log 2>&1; echo "full rc=$?"
echo done
