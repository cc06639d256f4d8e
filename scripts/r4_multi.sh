# Multi-GPU records of round 2 with the final binary: the driver's line (strong split of 16,384 economies, weak beside it)
set -x
P=gpurun_out/profiles; mkdir -p $P
nvidia-smi -L | wc -l
python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -2
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 > $P/r02_bench_${n}gpu_config3.json 2> gpurun_out/r4_bench_${n}gpu.err
echo "n=$n rc=$?"; python - <<PY
import json
d=json.load(open("$P/r02_bench_${n}gpu_config3.json"))
print(d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["economies_per_gpu"], d.get("also_measured"))
PY
done
