set -x
python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r2_t5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5.log
tail -5 gpurun_out/r2_t5.log
for c in 3 2 4 5; do
  python bench.py --config $c --steps 100 --warmup 5 > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err
  echo "config $c rc=$?"; head -c 900 gpurun_out/r2_bench_c$c.json; echo; tail -3 gpurun_out/r2_bench_c$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; head -c 600 gpurun_out/r2_bench_ref.json
