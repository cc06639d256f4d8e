set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_t6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t6.log
tail -5 gpurun_out/r2_t6.log
CATS_B200_LIB=build/libcats_prof.so python scripts/r2_mw_phases.py > gpurun_out/r2_mw_phases4.log 2>&1
grep -E "warps=(4|8)" gpurun_out/r2_mw_phases4.log
python scripts/r2_mw_perf.py > gpurun_out/r2_mw_perf4.log 2>&1
grep -E "solo|config2|config5|strong8" gpurun_out/r2_mw_perf4.log
python scripts/config1_perf.py > gpurun_out/r2_config1.log 2>&1; tail -5 gpurun_out/r2_config1.log
