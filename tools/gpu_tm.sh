# usage: bash tools/gpu_tm.sh <tag>: experiment build, tensor-memory kernel off / one / two positions per warp
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
XPB200_TM=1 XPB200_TM_NPW=1 timeout 120 python bench.py --config c2 --positions 20000 --no-cpu-baseline --no-cli > $D/small_tm1.log 2>&1; echo "small tm1 rc=$?"; tail -c 300 $D/small_tm1.log
XPB200_TM=1 XPB200_TM_NPW=2 timeout 120 python bench.py --config c2 --positions 20000 --no-cpu-baseline --no-cli > $D/small_tm2.log 2>&1; echo "small tm2 rc=$?"; tail -c 300 $D/small_tm2.log
for c in c2 c3 c4; do
  for v in "0 1" "1 1" "1 2"; do
    set -- $v
    XPB200_TM=$1 XPB200_TM_NPW=$2 timeout 300 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_${c}_tm$1_npw$2.log 2>&1
    echo "$c tm=$1 npw=$2 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_${c}_tm$1_npw$2.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_${c}_tm$1_npw$2.log) $(grep -o '"fit": [0-9.]*' $D/bench_${c}_tm$1_npw$2.log)"
  done
done
for v in "0 1" "1 1"; do
  set -- $v
  XPB200_TM=$1 XPB200_TM_NPW=$2 timeout 400 python bench.py --config c5 --positions 400000 --no-cpu-baseline --no-cli > $D/bench_c5_tm$1_npw$2.log 2>&1
  echo "c5 tm=$1 npw=$2 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_c5_tm$1_npw$2.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_c5_tm$1_npw$2.log) $(grep -o '"fit": [0-9.]*' $D/bench_c5_tm$1_npw$2.log)"
done
unset XPB200_LIB
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -5 $D/tests.log
