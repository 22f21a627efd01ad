set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu9.log
timeout 200 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_l.json 2>gpurun_out/bench_l.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_l.json').read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'k1', d['roofline']['kernel_ms'], 'mc', d['roofline_mc']['ms'], d['roofline_mc']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_l.log 2>&1
grep -E "mc_count|mc_emit|mc_totals" gpurun_out/launches_l.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -k2 | head -40
