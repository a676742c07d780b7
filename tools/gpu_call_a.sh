#!/bin/bash
# round-2 validation call: parity tests, slicing/conversion timing, bench line, DRAM traffic of the CRT Gram kernel
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gram_jobs.py -m gpu -x -q > gpurun_out/r2b_pytest_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_a.log
timeout 200 python tools/time_convert.py > gpurun_out/r2b_time_convert.json 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_a.json 2> gpurun_out/r2b_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
  --clock-control none -k regex:"k_gram_i8|k_slice_crt|k_gram_crt_final" -c 3 --csv --log-file gpurun_out/r2b_ncu_gram_crt.csv python tools/profile_gram.py 4096 65536 crt > gpurun_out/r2b_ncu_gram_crt.log 2>&1
tail -4 gpurun_out/r2b_pytest_a.log; cat gpurun_out/r2b_time_convert.json; head -c 600 gpurun_out/r2b_bench_a.json; echo; tail -5 gpurun_out/r2b_ncu_gram_crt.csv | cut -c1-300
