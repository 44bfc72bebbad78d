set -x
D=gpurun_out/$1
mkdir -p $D
timeout 600 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
timeout 300 python bench.py --config c2 --no-cpu-baseline > $D/bench_c2.log 2>&1
timeout 300 python bench.py --config c3 --no-cpu-baseline > $D/bench_c3.log 2>&1
timeout 300 python bench.py --config c4 --no-cpu-baseline > $D/bench_c4.log 2>&1
tail -3 $D/tests.log
