# usage: bash tools/gpu_ncu_tm.sh <tag> "<tm> <npw>"...: ncu --set full of the c2 single-warp-class fit kernel of the experiment build
D=gpurun_out/$1; shift
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for v in "$@"; do
  set -- $v
  XPB200_TM=$1 XPB200_TM_NPW=$2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xp_fit_tm|xp_fit_kernel<1" -c 1 -f -o $D/fit_c2_tm$1_npw$2 \
    python bench.py --config c2 --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_tm$1_npw$2.log 2>&1
  echo "tm $1 npw $2 rc=$?"
done
