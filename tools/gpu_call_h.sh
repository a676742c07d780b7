#!/bin/bash
# 8-GPU call: bench lines at N = 8 and N = 4 with the final defaults (residue planes, ring order, one push stream)
mkdir -p gpurun_out
for N in 8 4; do
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/r2h_scale_$N.err > gpurun_out/r2h_scale_$N.json
done
for N in 8 4; do head -c 330 gpurun_out/r2h_scale_$N.json; echo; tail -2 gpurun_out/r2h_scale_$N.err; done
