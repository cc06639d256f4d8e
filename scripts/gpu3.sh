set -x
python -m pytest tests/test_multiwarp_gpu.py tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/r2_t3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
CATS_B200_LIB=build/libcats_prof.so python scripts/r2_mw_phases.py > gpurun_out/r2_mw_phases2.log 2>&1
cat gpurun_out/r2_mw_phases2.log
python scripts/r2_mw_perf.py > gpurun_out/r2_mw_perf2.log 2>&1
cat gpurun_out/r2_mw_perf2.log
