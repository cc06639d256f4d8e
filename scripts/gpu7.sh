set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_t7.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7.log
tail -4 gpurun_out/r2_t7.log
python scripts/r2_mw_perf.py > gpurun_out/r2_mw_perf5.log 2>&1
grep -E "warps=1|warps=0" gpurun_out/r2_mw_perf5.log
python bench.py --steps 300 --warmup 10 > gpurun_out/r2_bench_c3b.json 2> gpurun_out/r2_bench_c3b.err; head -c 400 gpurun_out/r2_bench_c3b.json; echo
