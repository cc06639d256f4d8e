set -x
O=gpurun_out
# ncu: launch lists + the captures that depend on the final sources
NCU="ncu --clock-control none"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_plain_bench.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_ncu_bench.log 2>&1
python scripts/r2_ncu_target.py rollout 4096 10 > $O/r02_plain_rollout.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_rollout.csv python scripts/r2_ncu_target.py rollout 4096 10 > $O/r02_ncu_rollout.log 2>&1
full() {
  name=$1; rx=$2; skip=$3; shift 3
  python scripts/r2_ncu_target.py "$@" > $O/r02_plain_$name.log 2>&1 &&
  $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_r02_$name python scripts/r2_ncu_target.py "$@" > $O/r02_ncufull_$name.log 2>&1
  echo "$name rc=$?"
}
full c3 cats_period_kernel 1 period 16384 1 1000 100 20 1
full c2mw cats_period_mw_kernel 1 period 1024 1 1000 100 20 0
full tape cats_period_kernel 0 tape 1024
ls -la $O/prof_r02_*.ncu-rep
