#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c16
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "separate or sep or ragged or dropin or fp16 or forward_eval" > ${P}_pytest_sep.log 2>&1; echo "pytest(sep) rc=$?"; tail -5 ${P}_pytest_sep.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras --separate"
for v in sep4 nosep4 sep4 nosep4; do
  if [ $v = sep4 ]; then $B; else MCMIL_NO_SEP4=1 $B; fi 2>${P}_ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); r=b['roofline']
print('$v', 'value %.0f ms_per_step %.3f kernel_ms %.3f launches %d'%(b['value'], b['ms_per_step'], r['kernel_ms'], r['kernel_launches']))"
done > ${P}_ab.log 2>&1
cat ${P}_ab.log
tail -3 ${P}_ab_sep4.err
timeout 900 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
echo done
