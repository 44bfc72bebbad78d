# usage: bash tools/gpu_round.sh <tag> [configs...]: GPU tests, then short bench lines (no CPU baseline) of the configs
set -x
D=gpurun_out/$1
shift
mkdir -p $D
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $D/nvsmi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
for c in "$@"; do timeout 400 python bench.py --config $c --no-cpu-baseline > $D/bench_$c.log 2>&1; done
tail -5 $D/tests.log
