# usage: bash tools/gpu_lpt.sh <tag>: worklist ordered by expected sweeps (XPB200_LPT=1) or by read count alone (0), experiment build
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for c in c2 c5; do
  for t in 0 1 0 1; do
    extra=""; [ $c = c5 ] && extra="--positions 500000"
    XPB200_LPT=$t timeout 400 python bench.py --config $c $extra --no-cpu-baseline --no-cli > $D/bench_${c}_lpt$t.log 2>&1
    echo "$c lpt=$t rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_${c}_lpt$t.log | tr '\n' ' ') $(grep -o '"fit": [0-9.]*' $D/bench_${c}_lpt$t.log) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_${c}_lpt$t.log)"
  done
done
unset XPB200_LIB
timeout 900 python -m pytest tests -m gpu -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -3 $D/tests.log
