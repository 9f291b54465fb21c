B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:proj_tc -c 5 --csv --log-file gpurun_out/traffic_c2.csv $B > gpurun_out/ncu_t.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:proj_tc -c 10 --csv --log-file gpurun_out/traffic_c2_sep.csv $B --separate > gpurun_out/ncu_t2.log 2>&1
echo done
