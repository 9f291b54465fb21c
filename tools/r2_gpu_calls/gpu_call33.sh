#!/bin/bash
# throughput mode: PDL on/off x SMs reserved for the reduction kernels
mkdir -p gpurun_out
P=gpurun_out/r2c33
for pdl in on off; do
  if [ $pdl = on ]; then unset MCMIL_NO_PDL; else export MCMIL_NO_PDL=1; fi
  python tools/stream_probe.py 4,8 0,16,28 > ${P}_probe_pdl_$pdl.log 2>&1; echo "probe pdl=$pdl rc=$?"; cat ${P}_probe_pdl_$pdl.log
done
echo done
