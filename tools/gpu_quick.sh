# usage: bash tools/gpu_quick.sh <tag> [configs...]: GPU tests, then a bench line per config without the CPU / CLI legs
D=gpurun_out/$1
shift
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -4 $D/tests.log
for c in "$@"; do
  timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.log 2>&1
  echo "$c rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_$c.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_$c.log) $(grep -o '"worklist": [0-9.]*' $D/bench_$c.log) $(grep -o '"fit": [0-9.]*' $D/bench_$c.log)"
done
