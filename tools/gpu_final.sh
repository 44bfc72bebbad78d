# usage: bash tools/gpu_final.sh <tag>: GPU tests, smoke, bench lines of c2-c5 and the reference arm, ncu launch list and
# full captures of the fit and prefilter kernels at full c2 size (everything lands in gpurun_out/<tag>/)
D=gpurun_out/$1
mkdir -p $D
timeout 600 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
timeout 600 python bench.py > $D/bench_c2.log 2>&1
for c in c3 c4 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline > $D/bench_$c.log 2>&1; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:xp_fit_kernel -c 1 -f -o $D/fit_full \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $D/ncu_fit.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:xp_prefilter -c 1 -f -o $D/prefilter_full \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $D/ncu_pf.log 2>&1
tail -2 $D/tests.log; tail -1 $D/smoke.log
