#!/bin/bash
# 1 -> 8 GPU strong-scaling run of bench.py on one box; one JSON line per N in gpurun_out/scale_*.json
mkdir -p gpurun_out
MAXN=${1:-8}
for N in 1 2 4 8; do
  if [ $N -gt $MAXN ]; then break; fi
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 2>gpurun_out/scale_$N.err > gpurun_out/scale_$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/scale_$N.err > gpurun_out/scale_$N.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$N.json").read().strip().splitlines()[-1])
    print("N=$N", "ms_per_step=%.2f" % d["ms_per_step"], "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "launches", d["gpu_launches"])
except Exception as e:
    print("N=$N failed:", e); print(open("gpurun_out/scale_$N.err").read()[-1500:])
PY
done
