# usage: bash tools/gpu_chain2.sh <tag>: chunk plans of the chained host pipeline on c2 (experiment build)
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
line() { echo "$1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $2 | tr '\n' ' ')"; }
for cfg in "16 8" "4 8" "16 1" "16 32" "64 32"; do
  set -- $cfg
  XPB200_CHAIN_FIRST_DIV=$1 XPB200_CHAIN_LAST_DIV=$2 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_f$1_l$2.log 2>&1; line "c2 first=1/$1 last=1/$2" $D/bench_c2_f$1_l$2.log
done
XPB200_CHAIN_CHUNK_MB=16 XPB200_CHAIN_LAST_DIV=16 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_mb16.log 2>&1; line "c2 mb16 last=1/16" $D/bench_c2_mb16.log
XPB200_CHAIN_FREE_SMS=4 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_free4.log 2>&1; line "c2 free4" $D/bench_c2_free4.log
XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
tail -34 $D/trace_c2.log | cut -c1-150
