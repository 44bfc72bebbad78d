# usage: bash tools/gpu_chain5.sh <tag>: chained host pipeline - statistics / rows of finished chunks while later chunks are fitted (1) or after the launch (0) (experiment build)
D=gpurun_out/$1
mkdir -p $D
timeout 600 python -m pytest tests -m gpu -q -x > $D/tests.log 2>&1; echo "tests rc=$?"; tail -2 $D/tests.log
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
line() { echo "$1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $2 | tr '\n' ' ')"; }
for p in 0 1 0 1; do
  XPB200_CHAIN_POST=$p timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_post$p.log 2>&1; line "c2 post=$p" $D/bench_c2_post$p.log
done
for p in 0 1; do
  XPB200_CHAIN_POST=$p timeout 400 python bench.py --config c3 --no-cpu-baseline --no-cli > $D/bench_c3_post$p.log 2>&1; line "c3 post=$p" $D/bench_c3_post$p.log
done
XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
grep chained $D/trace_c2.log | tail -3
