# usage: bash tools/gpu_pf.sh <tag>: integer issue-rate microbenchmark, GPU tests, bench lines of c2 / c3 / c4 without the CPU / CLI legs
D=gpurun_out/$1
mkdir -p $D
[ -x build/exp/int_rate ] && timeout 120 build/exp/int_rate > $D/int_rate.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -5 $D/tests.log
for c in c2 c3 c4; do
  timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.log 2>&1
  echo "$c rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_$c.log | tr '\n' ' ') $(grep -o '"prefilter": [0-9.]*' $D/bench_$c.log) $(grep -o '"fit": [0-9.]*' $D/bench_$c.log) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_$c.log)"
done
cat $D/int_rate.txt
