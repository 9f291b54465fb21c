timeout 500 python tools/stress_gpu.py 0 80 > gpurun_out/stress.log 2>&1; echo "rc=$?" >> gpurun_out/stress.log; tail -4 gpurun_out/stress.log
