bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1
cat gpurun_out/tests_summary.txt
python bench.py > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
cat gpurun_out/bench_v4.json
python scripts/profile_step.py > gpurun_out/profile_step_plain.txt 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_dram_v2.csv python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/profile_step_plain.txt
