set -x
python -m pytest tests/test_multiwarp_gpu.py -x -q -m gpu > gpurun_out/r2_t2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2.log
tail -15 gpurun_out/r2_t2.log
python scripts/r2_mw_perf.py > gpurun_out/r2_mw_perf.log 2>&1
cat gpurun_out/r2_mw_perf.log
