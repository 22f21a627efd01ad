set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu8.log
timeout 200 python bench.py --steps 20 > gpurun_out/bench_k.json 2>gpurun_out/bench_k.err; tail -c 600 gpurun_out/bench_k.json
timeout 200 python bench.py --mode sf3d --steps 10 > gpurun_out/bench_k_sf3d.json 2>>gpurun_out/bench_k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_k.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lattice_tc_ta|mc_emit|mc_count|project_planes' -s 8 -c 4 -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_k2.log 2>&1
ls -la gpurun_out/prof_k.ncu-rep
