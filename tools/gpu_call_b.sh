#!/bin/bash
# 2-GPU call: full single-GPU parity suite, distributed parity, timelines, bench at N = 1, 2
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_b.log
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_timeline.py > gpurun_out/r2b_timeline_n2.log 2>&1
bash tools/scale.sh 2 > gpurun_out/r2b_scale2.log 2>&1
tail -4 gpurun_out/r2b_pytest_b.log; tail -3 gpurun_out/r2b_timeline_n2.log; cat gpurun_out/r2b_scale2.log
