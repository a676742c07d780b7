#!/bin/bash
# final single-GPU call of round 2: whole GPU suite, smoke, bench line, DRAM traffic + durations of the kernels of a bench step
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max \
  --clock-control none -k regex:"k_tile_pass|k_slice_crt|k_gram_i8|k_gram_crt_final8" -s 10 -c 5 --csv --log-file gpurun_out/r2g_ncu_step.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2g_ncu_step.log 2>&1
tail -3 gpurun_out/r2g_pytest.log; tail -2 gpurun_out/r2g_smoke.log; head -c 1500 gpurun_out/r2g_bench.json; echo; tail -3 gpurun_out/r2g_bench.err
