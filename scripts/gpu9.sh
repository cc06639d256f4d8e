set -x
python -m pytest tests/test_multiwarp_gpu.py tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/r2_t9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t9.log
tail -4 gpurun_out/r2_t9.log
python scripts/r2_pipe_perf.py > gpurun_out/r2_pipe_perf.log 2>&1; cat gpurun_out/r2_pipe_perf.log
