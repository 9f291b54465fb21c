timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_r22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r22.log; tail -12 gpurun_out/t_r22.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras"
$B > gpurun_out/ab_bias.json 2>&1
MCMIL_LIB_PATH=build/variants/prev.so $B > gpurun_out/ab_bias_prev.json 2>&1
$B > gpurun_out/ab_bias2.json 2>&1
MCMIL_LIB_PATH=build/variants/prev.so $B > gpurun_out/ab_bias_prev2.json 2>&1
timeout 300 python tools/stress_gpu.py 8 40 > gpurun_out/stress9.log 2>&1; tail -1 gpurun_out/stress9.log
