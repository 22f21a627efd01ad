#!/bin/bash
# One-call validation on a B200 box (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh'
# GPU parity tests, smoke(), the default bench line; everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('step_ms', round(d['ms_per_step'], 3), 'e2e_ms', round(d['e2e']['ms_per_step'], 3), 'K1_ms', round(d['roofline']['kernel_ms'], 3),
      'K1_frac', round(d['roofline']['frac'], 3), 'MC_ms', round(d['roofline_mc']['ms'], 3), 'MC_frac', round(d['roofline_mc']['frac'], 3),
      'cpu_pts_s', round(d['cpu_baseline']['value']))
PY
