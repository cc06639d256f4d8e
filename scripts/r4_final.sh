# Round-2 final records with the shipped binary, one gpurun call on one GPU: ncu summaries (written on the box, the bench
# kernel's traffic file tied to the source hash), then the bench lines of every configuration, the reference arm, the
# strict-mode rate, the adapter rate and the shape table.  Everything lands in gpurun_out/profiles/.
set -x
O=gpurun_out
P=$O/profiles
mkdir -p $P
bash scripts/r3_profile.sh > $O/r4_profile.log 2>&1
cp profiles/r02_ncu_*.md profiles/roofline_traffic.json $P/
for c in 3 2 4 5; do
  python bench.py --config $c --steps 300 --warmup 10 > $P/r02_bench_1gpu_config$c.json 2> $O/r4_bench_config$c.err
  echo "config $c rc=$?"; python - <<PY
import json
d=json.load(open("$P/r02_bench_1gpu_config$c.json"))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","vs_cpu_single_thread")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("traffic_stale"), d["step_ms"], d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["single_thread_value"], d["cpu_baseline"]["cores"], d["launch"])
PY
done
python bench.py --impl reference --steps 5 --warmup 1 > $P/r02_bench_reference_arm.json 2>/dev/null; head -c 300 $P/r02_bench_reference_arm.json; echo
python scripts/r2_strict_perf.py > $P/r02_strict_mode_perf.log 2>&1; cat $P/r02_strict_mode_perf.log
python scripts/config1_perf.py > $P/r02_config1_adapter.log 2>&1; tail -3 $P/r02_config1_adapter.log
(python scripts/r4_shape_sweep.py 148 512 1024 2048 4096 16384; python scripts/r4_c5_sweep.py) > $P/r02_shapes_perf.log 2>&1; cat $P/r02_shapes_perf.log
CATS_B200_LIB=build/libcats_prof.so python scripts/r2_mw_phases.py > $P/r02_mw_phase_cycles.log 2>&1; tail -8 $P/r02_mw_phase_cycles.log
