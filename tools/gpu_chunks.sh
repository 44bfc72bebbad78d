# usage: bash tools/gpu_chunks.sh <tag>: e2e of c2 over chunk size cap / growth / compute streams of the host pipeline (experiment build)
D=gpurun_out/$1
mkdir -p $D
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for v in "112 150 3" "112 200 3" "160 150 3" "224 150 3" "224 200 3" "80 150 3" "112 150 2" "112 130 3" "56 150 3"; do
  set -- $v
  XPB200_CHUNK_MB=$1 XPB200_CHUNK_GROWTH_PCT=$2 XPB200_STREAMS=$3 timeout 300 python bench.py --no-cpu-baseline --no-cli > $D/bench_$1_$2_$3.log 2>&1
  echo "chunk=$1 growth=$2 streams=$3 $(grep -o '"ms_per_step": [0-9.]*' $D/bench_$1_$2_$3.log | tr '\n' ' ')"
done
