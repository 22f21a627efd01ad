#!/bin/bash
# Scaling sweep on one box: default dp bench and the 512^3 x-slab sharded bench at N = 1,2,4,8.
# usage: tools/scale_run.sh <outdir> [max_gpus]
out=${1:-gpurun_out/scale}; maxn=${2:-8}; mkdir -p "$out"
port=29700
for mode in dp sharded; do
  for n in 1 2 4 8; do
    [ "$n" -gt "$maxn" ] && continue
    port=$((port+1))
    if [ "$n" = 1 ]; then
      timeout 600 python bench.py --gpus 1 --mode $mode --steps 10 --warmup 3 --no-cpu-baseline > "$out/${mode}_n$n.json" 2> "$out/${mode}_n$n.err"
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --mode $mode --steps 10 --warmup 3 --no-cpu-baseline > "$out/${mode}_n$n.json" 2> "$out/${mode}_n$n.err"
    fi
    python - "$out/${mode}_n$n.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], d["scaling"], d.get("mesh"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
