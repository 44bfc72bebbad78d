# usage: bash tools/gpu_tm2.sh <tag>: experiment build, selected tensor-memory variants per config
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
run() { # config tm npw [extra bench args]
  c=$1; tm=$2; npw=$3; shift; shift; shift
  XPB200_TM=$tm XPB200_TM_NPW=$npw timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli "$@" > $D/bench_${c}_tm${tm}_npw$npw.log 2>&1
  echo "$c tm=$tm npw=$npw rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_${c}_tm${tm}_npw$npw.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_${c}_tm${tm}_npw$npw.log) $(grep -o '"fit": [0-9.]*' $D/bench_${c}_tm${tm}_npw$npw.log)"
}
run c2 1 2; run c3 1 2; run c4 1 2
XPB200_TM_GT=0 run c2 1 2
XPB200_TM=1 XPB200_TM_NPW=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xp_fit_tm" -c 1 -f -o $D/fit_c2_tm1_npw2 \
    python bench.py --config c2 --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_tm1_npw2.log 2>&1
