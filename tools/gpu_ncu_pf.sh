# usage: bash tools/gpu_ncu_pf.sh <tag>: one ncu --set full capture of the u16 prefilter at full c2 size
D=gpurun_out/$1
mkdir -p $D
timeout 900 ncu --set full --import-source on --clock-control none -k regex:xp_prefilter_u16 -c 1 -f -o $D/prefilter_u16 \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_pf.log 2>&1
echo "rc=$?"; tail -3 $D/ncu_pf.log
