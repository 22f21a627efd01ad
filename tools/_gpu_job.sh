set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu11.log
timeout 200 python bench.py --steps 20 > gpurun_out/bench_m.json 2>gpurun_out/bench_m.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_m.json').read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'k1', d['roofline']['kernel_ms'], d['roofline']['frac'], 'mc', d['roofline_mc']['ms'], d['roofline_mc']['frac'], d['cpu_baseline']['value'])"
timeout 100 python tools/bench_lattice.py 512 3 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
