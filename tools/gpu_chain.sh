# usage: bash tools/gpu_chain.sh <tag>: the chained host pipeline - GPU tests with the shipped library, then e2e of c2 / c3
# with the experiment build over free-SM counts and chunk sizes (XPB200_CHAIN=0: chunk by chunk launches)
D=gpurun_out/$1
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q -x > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -5 $D/tests.log
line() { echo "$1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $2 | tr '\n' ' ') $(grep -o '"sweeps_mean": [0-9.]*' $2)"; }
timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_ship.log 2>&1; line "c2 shipped" $D/bench_c2_ship.log
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
XPB200_CHAIN=0 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_chain0.log 2>&1; line "c2 chain=0" $D/bench_c2_chain0.log
for f in 4 6 8 12; do
  XPB200_CHAIN_FREE_SMS=$f timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_free$f.log 2>&1; line "c2 free=$f" $D/bench_c2_free$f.log
done
for mb in 16 64; do
  XPB200_CHAIN_CHUNK_MB=$mb timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_mb$mb.log 2>&1; line "c2 mb=$mb" $D/bench_c2_mb$mb.log
done
XPB200_CHAIN=0 timeout 400 python bench.py --config c3 --no-cpu-baseline --no-cli > $D/bench_c3_chain0.log 2>&1; line "c3 chain=0" $D/bench_c3_chain0.log
timeout 400 python bench.py --config c3 --no-cpu-baseline --no-cli > $D/bench_c3_chain1.log 2>&1; line "c3 chain=1" $D/bench_c3_chain1.log
XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
