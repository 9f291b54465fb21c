python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python bench.py --separate --no-cpu > gpurun_out/bench_full_sep.json 2> gpurun_out/bench_full_sep.err
B="python bench.py --steps 2 --warmup 1 --bags-per-step 32 --no-e2e --no-cpu --no-extras"
$B > gpurun_out/plain_r1.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v2.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"proj_tc|softmax_rows|welford_cols" -c 6 -o gpurun_out/prof_r1_v2 -f $B > gpurun_out/ncu_f.log 2>&1
echo done
