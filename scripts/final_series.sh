# The evidence series of a round state (run through gpurun): every -m gpu test file, smoke(), every bench.py config and the reference arm,
# the ncu launch list of one step (-> scripts/summarize_ncu.py traffic) and the in-graph timeline.  Outputs land in gpurun_out/.
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1
cat gpurun_out/tests_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; echo "bench rc $?"
python bench.py --impl reference > gpurun_out/bench_ref_v9.json 2> gpurun_out/bench_ref_v9.err; echo "ref rc $?"
python bench.py --config clip > gpurun_out/bench_clip_v7.json 2> gpurun_out/bench_clip_v7.err; echo "clip rc $?"
python bench.py --config crowded > gpurun_out/bench_crowded_v7.json 2> gpurun_out/bench_crowded_v7.err; echo "crowded rc $?"
python - <<'P'
import json
for n in ["bench_v9","bench_ref_v9","bench_clip_v7","bench_crowded_v7"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%n)); print(n, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"))
    except Exception as e: print(n, "ERR", e)
P
python scripts/profile_step.py > gpurun_out/profile_step_plain.txt 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_dram_v7.csv python scripts/profile_step.py > gpurun_out/ncu_launches7.log 2>&1
tail -1 gpurun_out/profile_step_plain.txt; tail -1 gpurun_out/ncu_launches7.log
python scripts/step_timeline.py > gpurun_out/timeline_v5.txt 2>&1
