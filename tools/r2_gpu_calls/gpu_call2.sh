#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c2
for path in auto split; do
  for skip in none reduce proj; do
    if [ $skip = none ]; then python tools/single_bag_probe.py 300 $path; else MCMIL_EXP_SKIP=$skip python tools/single_bag_probe.py 300 $path; fi 2>&1 | sed "s/^/skip=$skip /"
  done
done > ${P}_single.log 2>&1
cat ${P}_single.log
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 ${P}_pytest.log
python bench.py --steps 20 --warmup 5 --no-configs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
python tools/single_bag_probe.py 10 auto > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused -s 6 -c 1 -o ${P}_fused python tools/single_bag_probe.py 10 auto > ${P}_ncu_fused.log 2>&1
python tools/single_bag_probe.py 10 split > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'softmax|welford' -s 12 -c 2 -o ${P}_split python tools/single_bag_probe.py 10 split > ${P}_ncu_split.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
