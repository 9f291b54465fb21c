#!/bin/bash
# timing-only variants of the (reverted) 4-CTA-cluster separate-attention kernel: how much of its time are the remote stores?
mkdir -p gpurun_out
P=gpurun_out/r2c18
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --separate"
for v in base sep4 sep4_noremote sep4_noremote_cta; do
  if [ $v = base ]; then $B; else MCMIL_LIB_PATH=build/variants/$v.so $B; fi 2>${P}_ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); r=b['roofline']
print('$v', 'value %.0f ms_per_step %.3f kernel_ms %.3f launches %d'%(b['value'], b['ms_per_step'], r['kernel_ms'], r['kernel_launches']))"
done > ${P}_ab.log 2>&1
cat ${P}_ab.log
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
echo done
