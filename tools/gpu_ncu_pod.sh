# usage: bash tools/gpu_ncu_pod.sh <tag> <pod values...>: ncu --set full of the c2 fit kernel of the experiment build per XPB200_POD value
D=gpurun_out/$1; shift
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for pod in "$@"; do
  XPB200_POD=$pod timeout 600 ncu --set full --import-source on --clock-control none -k regex:xp_fit -c 1 -f -o $D/fit_c2_pod$pod \
    python bench.py --config c2 --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_pod$pod.log 2>&1
  echo "pod $pod rc=$?"
done
