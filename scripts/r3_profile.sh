# Final round-2 ncu captures with the shipped binary (one gpurun call, one GPU).  Each command runs plain first, then
# under ncu; the summaries (profiles/r02_*.md) are generated ON the box with the exact mangled kernel names and copied
# to gpurun_out/profiles/ (only gpurun_out/ travels back); the reports are deleted except the config-3 one.
set -x
O=gpurun_out
mkdir -p $O/profiles
NCU="ncu --clock-control none"
SO=$PWD/r_mabm_b200/libcats_b200.so
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_plain_bench.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/profiles/r02_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_ncu_bench.log 2>&1
python scripts/r2_ncu_target.py rollout 4096 10 > $O/r02_plain_rollout.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/profiles/r02_launches_rollout.csv python scripts/r2_ncu_target.py rollout 4096 10 > $O/r02_ncu_rollout.log 2>&1
full() {  # name, kernel regex, skip, target args...
  name=$1; rx=$2; skip=$3; shift 3
  python scripts/r2_ncu_target.py "$@" > $O/r02_plain_$name.log 2>&1 &&
  $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_r02_$name python scripts/r2_ncu_target.py "$@" > $O/r02_ncufull_$name.log 2>&1
  echo "$name rc=$?"
}
md() { python scripts/make_profile_md.py "$@" > $O/md_$6.log 2>&1 || tail -3 $O/md_$6.log; }
full c3 cats_period_kernel 1 period 16384 1 1000 100 20 1
python scripts/make_profile_md.py $O/prof_r02_c3.ncu-rep $SO auto 16384 r02 > $O/md_c3.log 2>&1 || tail -3 $O/md_c3.log
full c2mw cats_period_mw_kernel 1 period 1024 1 1000 100 20 0
md $O/prof_r02_c2mw.ncu-rep $SO auto 1024 r02 mw_kernel_config2 "1,024 default-size economies (BASELINE config 2), ONE period, after the burn-in: the multi-warp kernel, one block of 4 warps per economy with the pipelined goods market (chosen by the library)."
rm -f $O/prof_r02_c2mw.ncu-rep
full c5mw cats_period_mw_kernel 1 period 256 1 10000 1000 200 0
md $O/prof_r02_c5mw.ncu-rep $SO auto 256 r02 mw_kernel_config5 "256 economies of 10x size (BASELINE config 5), ONE period, after 100 burn-in periods: the multi-warp kernel, one block of 8 warps per economy."
rm -f $O/prof_r02_c5mw.ncu-rep
full c5one cats_period_kernel 1 period 256 1 10000 1000 200 1
md $O/prof_r02_c5one.ncu-rep $SO auto 256 r02 one_warp_kernel_config5 "256 economies of 10x size, ONE period: the one-warp-per-economy kernel (generic layout, MODE 1) forced with cats_debug_set_warps(1), for comparison with the multi-warp kernel."
rm -f $O/prof_r02_c5one.ncu-rep
full tape cats_period_kernel 0 tape 1024
md $O/prof_r02_tape.ncu-rep $SO auto 1024 r02 tape_kernel "1,024 default-size economies, ONE period driven by a FOREIGN replay tape (numpy draws; MODE 0: the kernel with every feature compiled in -- tape reads, debug stop, debug image), 8 economies per block.  Adds 56 KB of tape per economy-period to the traffic."
rm -f $O/prof_r02_tape.ncu-rep
full q cats_q_kernel 5 rollout 4096 10
md $O/prof_r02_q.ncu-rep $SO auto 8192 r02 q_kernel "config 4 rollout step: 4,096 economies x 2 agents = 8,192 (economy, agent) pairs, one warp each: TD(0) update of one cell + epsilon-greedy choice over the 121-cell row Q[s'] (11 x 11 doubles = 968 B read per pair; the 11^4-cell table per pair is 117 KB, 960 MB in all, touched sparsely)." "(economy, agent) pair"
rm -f $O/prof_r02_q.ncu-rep
full init cats_init_kernel 0 period 16384 1 1000 100 20 1
md $O/prof_r02_init.ncu-rep $SO auto 16384 r02 init_kernel "cats_init: 16,384 default-size economy blobs (37,760 B each, 619 MB) zeroed and initialised; one block of 256 threads per economy.  A pure HBM-write kernel: algorithmic bytes = 1 x blob." "economy"
rm -f $O/prof_r02_init.ncu-rep
cp profiles/r02_ncu_*.md profiles/roofline_traffic.json $O/profiles/
ls -la $O/profiles $O/*.ncu-rep
grep -h "rc=\|rror" $O/md_*.log | head
