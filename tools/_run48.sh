B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 100 $B > gpurun_out/y_base.json 2>&1
for v in noepi ponly mmaonly nomask; do MCMIL_LIB_PATH=build/variants/$v.so timeout 100 $B > gpurun_out/y_$v.json 2>&1; done
