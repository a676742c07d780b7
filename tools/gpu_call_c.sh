#!/bin/bash
# 2-GPU call: distributed parity, N=2 timelines (residue and digit exchange), conversion timing, launch list of a bench step
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_gram_jobs.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
for x in residues digits; do
B200SQ_DIST_CRT_EXCHANGE=$x timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dist_timeline.py > gpurun_out/r2c_timeline_n2_$x.log 2>&1
done
timeout 200 python tools/time_convert.py > gpurun_out/r2c_time_convert.json 2>&1
ROWS=2048 timeout 200 python tools/time_convert.py >> gpurun_out/r2c_time_convert.json 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches_c3.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c_ncu_bench.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; tail -2 gpurun_out/r2c_timeline_n2_residues.log; tail -2 gpurun_out/r2c_timeline_n2_digits.log; cat gpurun_out/r2c_time_convert.json
