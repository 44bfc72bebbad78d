# usage: bash tools/gpu_ncu_fit.sh <tag> : ncu --set full captures of the fit kernel on c3 (100k positions) and c2 (200k positions)
D=gpurun_out/$1
mkdir -p $D
for c in c3:100000 c2:200000; do
  cfg=${c%%:*}; np=${c##*:}
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:xp_fit_kernel -c 1 -f -o $D/fit_$cfg \
    python bench.py --config $cfg --positions $np --steps 1 --warmup 0 --no-cpu-baseline > $D/ncu_$cfg.log 2>&1
done
ls -la $D
