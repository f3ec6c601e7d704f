python scripts/profile_step.py > gpurun_out/profile_step_plain.txt 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_dram_v3.csv python scripts/profile_step.py > gpurun_out/ncu_launches3.log 2>&1
tail -1 gpurun_out/profile_step_plain.txt; tail -1 gpurun_out/ncu_launches3.log
ITERS=4 python scripts/conv_bench.py "reid.l1" > gpurun_out/cb_plain.txt 2>&1 && \
ITERS=4 ncu --set full --clock-control none --import-source on -k regex:conv_win_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2_conv_win_reid_l1 python scripts/conv_bench.py "reid.l1" > gpurun_out/ncu_full1.log 2>&1
tail -2 gpurun_out/ncu_full1.log
python scripts/k1_bench.py 8 6 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:preprocess_pairs_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_k1_pairs python scripts/k1_bench.py 64 6 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
ncu --set full --clock-control none --import-source on -k regex:assoc_kernel --launch-skip 20 -c 1 -f -o gpurun_out/r2_assoc python scripts/crowded_bench.py 64 17 30 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
