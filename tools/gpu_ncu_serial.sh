# usage: bash tools/gpu_ncu_serial.sh <tag>: bench.py with every kernel serialised by ncu (no launch filter, no count): the chained
# host pipeline must time out once, fall back to chunk-by-chunk launches and finish
D=gpurun_out/$1
mkdir -p $D
SECONDS=0; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_all.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cli > $D/ncu_all.log 2> $D/ncu_all.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' $D/ncu_all.log | tr '\n' ' '; echo "wall ${SECONDS}s"; wc -l $D/launches_all.csv; rm -f $D/launches_all.csv
