bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1
cat gpurun_out/tests_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench rc $?"
python bench.py --impl reference > gpurun_out/bench_ref_v6.json 2> gpurun_out/bench_ref_v6.err; echo "ref rc $?"
python bench.py --config clip > gpurun_out/bench_clip_v4.json 2> gpurun_out/bench_clip_v4.err; echo "clip rc $?"
python bench.py --config crowded > gpurun_out/bench_crowded_v4.json 2> gpurun_out/bench_crowded_v4.err; echo "crowded rc $?"
python - <<'P'
import json
for n in ["bench_v6","bench_ref_v6","bench_clip_v4","bench_crowded_v4"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%n)); print(n, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"))
    except Exception as e: print(n, "ERR", e)
P
python scripts/profile_step.py > gpurun_out/profile_step_plain.txt 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_dram_v4.csv python scripts/profile_step.py > gpurun_out/ncu_launches4.log 2>&1
tail -1 gpurun_out/profile_step_plain.txt; tail -1 gpurun_out/ncu_launches4.log
AICAM_BENCH_PAD=2 ITERS=4 python scripts/conv_bench.py "reid.l1" > gpurun_out/cb_plain.txt 2>&1 && \
AICAM_BENCH_PAD=2 ITERS=4 ncu --set full --clock-control none --import-source on -k regex:conv_win_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2_conv_win_reid_l1_ws python scripts/conv_bench.py "reid.l1" > gpurun_out/ncu_full4.log 2>&1
tail -2 gpurun_out/ncu_full4.log
python scripts/step_timeline.py > gpurun_out/timeline_v2.txt 2>&1
ls -la gpurun_out/*.ncu-rep
