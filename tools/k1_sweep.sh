#!/bin/bash
# Dev sweep of the K1 variants on one B200: parity tests of the lattice kernel, then timing per variant.
#   gpurun --timeout 600 -- 'bash tools/k1_sweep.sh'
mkdir -p gpurun_out
L=gpurun_out/k1_sweep.log
: > $L
timeout 300 python -m pytest tests/test_gpu_field.py -x -q -k "lattice or full_size" 2>&1 | tail -3 >> $L
for v in "SMB_TC_VARIANT=ta" "SMB_TC_VARIANT=pair SMB_TC_POLY=0" "SMB_TC_POLY=0 SMB_TC_WAITNS=200" "SMB_TC_POLY=0 SMB_TC_WAITNS=20000" "SMB_TC_POLY=1" "SMB_TC_POLY=2" "SMB_TC_POLY=3"; do
  echo "== $v" >> $L
  env $v timeout 120 python tools/bench_lattice.py 256 10 2>&1 | tail -1 >> $L
done
env SMB_TC_POLY=0 timeout 120 python tools/bench_lattice.py 512 5 2>&1 | tail -1 >> $L
cat $L
