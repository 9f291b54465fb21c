#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c17
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --separate"
for v in base s4_arrcta s4_waitcta s4_fencecta s4_all; do
  if [ $v = base ]; then $B; else MCMIL_LIB_PATH=build/variants/$v.so $B; fi 2>${P}_ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); r=b['roofline']
print('$v', 'value %.0f ms_per_step %.3f kernel_ms %.3f launches %d'%(b['value'], b['ms_per_step'], r['kernel_ms'], r['kernel_launches']))"
done > ${P}_ab.log 2>&1
cat ${P}_ab.log
MCMIL_LIB_PATH=build/variants/s4_all.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "separate or sep or ragged or dropin or fp16 or forward_eval" > ${P}_pytest_all.log 2>&1; echo "pytest(s4_all) rc=$?"; tail -3 ${P}_pytest_all.log
echo done
