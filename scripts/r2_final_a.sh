set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r2_tfinal.log 2>&1; echo "rc=$?" >> $O/r2_tfinal.log; tail -3 $O/r2_tfinal.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; tail -1 $O/r2_smoke.log
for c in 3 2 4 5; do
  python bench.py --config $c --steps 300 --warmup 10 > $O/r02_bench_config$c.json 2> $O/r02_bench_config$c.err
  echo "config $c rc=$?"; python - <<PY
import json
d=json.load(open("$O/r02_bench_config$c.json"))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","vs_cpu_single_thread")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("traffic_stale"), d["step_ms"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["single_thread_value"], d["cpu_baseline"]["cores"], d["launch"])
PY
done
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.json 2>/dev/null; head -c 200 $O/r02_bench_reference_arm.json; echo
python scripts/r2_strict_perf.py > $O/r2_strict.log 2>&1; cat $O/r2_strict.log
python scripts/config1_perf.py > $O/r2_config1.log 2>&1; tail -3 $O/r2_config1.log
python scripts/r2_pipe_perf.py > $O/r2_shapes_final.log 2>&1; cat $O/r2_shapes_final.log
