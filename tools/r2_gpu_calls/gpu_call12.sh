#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c12
MCMIL_LIB_PATH=build/variants/batch4.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or ragged or config2" > ${P}_pytest_batch4.log 2>&1; echo "pytest(batch4) rc=$?"; tail -2 ${P}_pytest_batch4.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras"
for v in base batch4 base batch4 base batch4; do
  if [ $v = base ]; then $B; else MCMIL_LIB_PATH=build/variants/$v.so $B; fi 2>${P}_ab_$v.err | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); r=b['roofline']
print('$v', 'kernel_ms %.3f frac_burst %.3f value %.0f clocks %s'%(r['kernel_ms'], r['frac_of_burst'], b['value'], b['clocks']['sm_mhz']))"
done > ${P}_ab.log 2>&1
cat ${P}_ab.log
MCMIL_LIB_PATH=build/variants/batch4.so python tools/single_bag_probe.py 300 x graph 2>&1 | sed 's/^/batch4 /'
python tools/single_bag_probe.py 300 x graph 2>&1 | sed 's/^/base /'
echo done
