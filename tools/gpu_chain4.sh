# usage: bash tools/gpu_chain4.sh <tag>: chained host pipeline - small arrays copied per chunk (XPB200_CHAIN_META_MB=1) or per group of chunks, chunk sizes (experiment build)
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
line() { echo "$1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $2 | tr '\n' ' ')"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chained or rows_equals or understated" 2>&1 | tail -2
for cfg in "1 32" "128 32" "128 16" "128 8" "256 16" "64 16"; do
  set -- $cfg
  XPB200_CHAIN_META_MB=$1 XPB200_CHAIN_CHUNK_MB=$2 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_meta$1_mb$2.log 2>&1; line "c2 meta=$1 chunk=$2" $D/bench_c2_meta$1_mb$2.log
done
XPB200_CHAIN_CHUNK_MB=16 XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
grep chained $D/trace_c2.log | tail -2
