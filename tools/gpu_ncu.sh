# usage: bash tools/gpu_ncu.sh <tag> <config> <positions|0> <kernel regex> <count> <name>: one ncu --set full capture
D=gpurun_out/$1; cfg=$2; np=$3; rx=$4; cnt=$5; name=$6
mkdir -p $D
pos=""; [ "$np" != "0" ] && pos="--positions $np"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$rx -c $cnt -f -o $D/$name \
  python bench.py --config $cfg $pos --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_$name.log 2>&1
echo "rc=$?" >> $D/ncu_$name.log
