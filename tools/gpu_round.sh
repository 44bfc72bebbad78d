# usage: bash tools/gpu_round.sh <tag> [configs...]: GPU tests, then bench lines of the configs (first one with all legs)
set -x
D=gpurun_out/$1
shift
mkdir -p $D
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $D/nvsmi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
first=1
for c in "$@"; do
  if [ $first = 1 ]; then timeout 900 python bench.py --config $c > $D/bench_$c.log 2>&1; first=0
  else timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.log 2>&1; fi
done
tail -5 $D/tests.log
