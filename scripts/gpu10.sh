set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_t10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t10.log
tail -4 gpurun_out/r2_t10.log
python scripts/r2_pipe_perf.py > gpurun_out/r2_pipe_perf2.log 2>&1; grep -E "warps=1|warps=0" gpurun_out/r2_pipe_perf2.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_c3c.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3c.json')); print(d['value'], d['step_ms'], d['roofline']['frac'])"
