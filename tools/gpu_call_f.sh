#!/bin/bash
# 8-GPU call: block order and push-stream variants on 8 and 4 ranks (timelines + parity), then the bench at N = 4, 8
mkdir -p gpurun_out
for n in 8 4; do
VARIANTS="1:full:auto:ring:1,1:full:auto:narrow:1,1:full:auto:narrow:2,1:full:auto:ring:2" timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/dist_timeline.py > gpurun_out/r2f_timeline_n$n.log 2>&1
done
for N in 4 8; do
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/r2f_scale_$N.err > gpurun_out/r2f_scale_$N.json
done
grep "^#" gpurun_out/r2f_timeline_n8.log; grep "^#" gpurun_out/r2f_timeline_n4.log; tail -3 gpurun_out/r2f_timeline_n4.log | cut -c1-300; for N in 4 8; do head -c 330 gpurun_out/r2f_scale_$N.json; echo; done
