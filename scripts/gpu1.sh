set -x
nvidia-smi -L
python -m pytest tests/test_replay_gpu.py tests/test_strict_gpu.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log
tail -15 gpurun_out/r2_t1.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2_tall.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tall.log
tail -5 gpurun_out/r2_tall.log
CATS_B200_LIB=build/libcats_prof.so python scripts/r2_phases.py > gpurun_out/r2_phases.log 2>&1
cat gpurun_out/r2_phases.log
python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
cat gpurun_out/r2_bench0.json | head -c 1500
