#!/bin/bash
# full ncu captures of the two reduction kernels at config 4 (one bag, N=16384, T=1000)
mkdir -p gpurun_out
P=gpurun_out/r2c21
CMD="python bench.py --workload config4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras"
$CMD > ${P}_plain.json 2> ${P}_plain.err; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'welford_cols' -s 2 -c 1 -o ${P}_cols $CMD > ${P}_ncu_cols.log 2>&1; echo "ncu cols rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'softmax_rows' -s 2 -c 1 -o ${P}_rows $CMD > ${P}_ncu_rows.log 2>&1; echo "ncu rows rc=$?"
echo done
