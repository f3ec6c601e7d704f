bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1
cat gpurun_out/tests_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench rc $?"
python bench.py --impl reference > gpurun_out/bench_ref_v5.json 2> gpurun_out/bench_ref_v5.err; echo "ref rc $?"
python bench.py --config clip > gpurun_out/bench_clip_v2.json 2> gpurun_out/bench_clip_v2.err; echo "clip rc $?"
python bench.py --config crowded > gpurun_out/bench_crowded_v3.json 2> gpurun_out/bench_crowded_v3.err; echo "crowded rc $?"
python bench.py --config sweep > gpurun_out/bench_sweep_v2.json 2> gpurun_out/bench_sweep_v2.err; echo "sweep rc $?"
python scripts/k1_bench.py > gpurun_out/k1_bench.txt 2>&1
python scripts/crowded_bench.py 16 300 140 > gpurun_out/crowded_tracker.txt 2>&1
python - <<'P'
import json
for n in ["bench_v5","bench_ref_v5","bench_clip_v2","bench_crowded_v3","bench_sweep_v2"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%n)); print(n, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(n, "ERR", e)
P
