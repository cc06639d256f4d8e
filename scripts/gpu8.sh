set -x
nvidia-smi -L
python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r2_t8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8.log
tail -4 gpurun_out/r2_t8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "rc=$?"; head -c 600 gpurun_out/r2_bench_2gpu.json; echo; grep -o '"also_measured".*' gpurun_out/r2_bench_2gpu.json | head -c 400; echo; tail -3 gpurun_out/r2_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/r2_bench_2gpu_ref.json 2>> gpurun_out/r2_bench_2gpu.err; head -c 300 gpurun_out/r2_bench_2gpu_ref.json
