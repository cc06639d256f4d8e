# one --set full capture of the config-3 launch (report kept under gpurun_out/ for line-level analysis)
set -x
O=gpurun_out
python scripts/r2_ncu_target.py period 16384 1 1000 100 20 1 > $O/r03_plain_c3.log 2>&1 &&
ncu --clock-control none --set full --import-source on -k regex:cats_period_kernel -s 1 -c 1 -f -o $O/prof_r03_c3 python scripts/r2_ncu_target.py period 16384 1 1000 100 20 1 > $O/r03_ncufull_c3.log 2>&1
echo "c3 rc=$?"
ls -la $O/*.ncu-rep
