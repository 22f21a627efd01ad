#!/bin/bash
# Dev sweep (needs the developer build) of the number of activation PAIRS per step that take the FMA-pipe polynomial.
#   python -m sculptmate_b200.build --dev && gpurun --timeout 600 -- 'bash tools/k1_poly_sweep.sh'
mkdir -p gpurun_out
L=gpurun_out/k1_poly_sweep.log
: > $L
for n in ${K1_POLY_LIST:-0 1 2 3 4 5 6 8 10 12}; do
  echo "== SMB_TC_TA_POLY=$n (pairs of 32)" >> $L
  env SMB_TC_TA_POLY=$n timeout 120 python tools/bench_lattice.py 256 10 2>&1 | tail -1 >> $L
done
timeout 300 python -m pytest tests/test_gpu_field.py -x -q -k "lattice or full_size" 2>&1 | tail -2 >> $L
cat $L
