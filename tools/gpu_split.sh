# usage: bash tools/gpu_split.sh <tag>: split between the single-warp tensor-memory class and the 4-warp teams (experiment build)
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
run() { # config split [extra]
  c=$1; sp=$2; shift; shift
  XPB200_TM_SPLIT=$sp timeout 400 python bench.py --config $c --no-cpu-baseline --no-cli "$@" > $D/bench_${c}_split$sp.log 2>&1
  echo "$c split=$sp rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_${c}_split$sp.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_${c}_split$sp.log) $(grep -o '"fit": [0-9.]*' $D/bench_${c}_split$sp.log) $(grep -o '"warps_per_cta": [0-9]*' $D/bench_${c}_split$sp.log)"
}
for sp in 1024 1152 1280 1536 2048; do run c5 $sp --positions 400000; done
for sp in 1024 1152 1536 2048; do run c4 $sp; done
