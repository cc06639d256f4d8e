set -x
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r02_bench_${n}gpu_config3.json 2> gpurun_out/r02_bench_${n}gpu.err
echo "n=$n rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${n}gpu_config3.json"))
print(d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["economies_per_gpu"], d["config"]["l2_policy"][:60], d.get("also_measured"))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference > gpurun_out/r02_bench_8gpu_ref.json 2>> gpurun_out/r02_bench_8gpu.err; head -c 200 gpurun_out/r02_bench_8gpu_ref.json
