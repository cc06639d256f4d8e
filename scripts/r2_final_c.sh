set -x
O=gpurun_out
NCU="ncu --clock-control none"
full() {
  name=$1; rx=$2; skip=$3; shift 3
  python scripts/r2_ncu_target.py "$@" > $O/r02_plain_$name.log 2>&1 &&
  $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_r02_$name python scripts/r2_ncu_target.py "$@" > $O/r02_ncufull_$name.log 2>&1
  echo "$name rc=$?"
}
full c5mw cats_period_mw_kernel 1 period 256 1 10000 1000 200 0
full c5one cats_period_kernel 1 period 256 1 10000 1000 200 1
ls -la $O/prof_r02_*.ncu-rep
