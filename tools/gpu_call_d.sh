#!/bin/bash
# 2-GPU call: distributed parity incl. CRT exchange variants, N=2 timelines (whole-block and sub-block pushes), bench N=1,2
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -s > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
for sub in 1 4; do
B200SQ_DIST_SUBBLOCKS=$sub timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_timeline.py > gpurun_out/r2d_timeline_n2_sub$sub.log 2>&1
done
bash tools/scale.sh 2 > gpurun_out/r2d_scale2.log 2>&1
tail -4 gpurun_out/r2d_pytest.log; tail -2 gpurun_out/r2d_timeline_n2_sub1.log; tail -2 gpurun_out/r2d_timeline_n2_sub4.log; cat gpurun_out/r2d_scale2.log
