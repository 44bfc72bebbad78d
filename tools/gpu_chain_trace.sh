# usage: bash tools/gpu_chain_trace.sh <tag> [env assignments...]: per-chunk timeline of the chained host pipeline on c2 (experiment build)
D=gpurun_out/$1
shift
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for kv in "$@"; do export "$kv"; done
XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
echo "rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_c2_trace.log | tr '\n' ' ')"
tail -32 $D/trace_c2.log
