# usage: bash tools/gpu_chain3.sh <tag>: chained host pipeline with / without the late helper launch on the freed SMs (experiment build)
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
line() { echo "$1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $2 | tr '\n' ' ')"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chained or rows_equals or understated" 2>&1 | tail -2
for h in 0 1; do
  for c in c2 c3; do
    XPB200_CHAIN_HELPER=$h timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_${c}_helper$h.log 2>&1; line "$c helper=$h" $D/bench_${c}_helper$h.log
  done
  XPB200_CHAIN_HELPER=$h timeout 400 python bench.py --config c5 --positions 400000 --no-cpu-baseline --no-cli > $D/bench_c5_helper$h.log 2>&1; line "c5 400k helper=$h" $D/bench_c5_helper$h.log
done
