# usage: bash tools/gpu_round2.sh <tag>: GPU tests, full c2 bench line, c3-c5 lines, prefilter occupancy experiment
D=gpurun_out/$1
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -3 $D/tests.log
timeout 900 python bench.py > $D/bench_c2.log 2>&1; echo "c2 rc=$?"
for c in c3 c4 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.log 2>&1; echo "$c rc=$?"; done
XPB200_LIB=$PWD/build/exp/libxp_pf8.so timeout 300 python bench.py --no-cpu-baseline --no-cli > $D/bench_c2_pf8.log 2>&1
grep -o '"prefilter": [0-9.]*' $D/bench_c2.log $D/bench_c2_pf8.log
