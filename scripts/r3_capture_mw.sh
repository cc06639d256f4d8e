# --set full captures of the multi-warp kernel at config 5 and config 2 (reports kept under gpurun_out/)
set -x
O=gpurun_out
full() {  # name, kernel regex, skip, target args...
  name=$1; rx=$2; skip=$3; shift 3
  python scripts/r2_ncu_target.py "$@" > $O/r03_plain_$name.log 2>&1 &&
  ncu --clock-control none --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_r03_$name python scripts/r2_ncu_target.py "$@" > $O/r03_ncufull_$name.log 2>&1
  echo "$name rc=$?"
}
full c5mw cats_period_mw_kernel 1 period 256 1 10000 1000 200 0
full c2mw cats_period_mw_kernel 1 period 1024 1 1000 100 20 0
