# usage: bash tools/gpu_multi.sh <tag> <N> [bench args]: the 2-rank CLI test and bench.py on N GPUs under torchrun
set -x
D=gpurun_out/$1; N=$2; shift; shift
mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "two_ranks" > $D/tests_multi.log 2>&1; echo "rc=$?" >> $D/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N "$@" > $D/bench_${N}gpu.log 2>&1; echo "rc=$?" >> $D/bench_${N}gpu.log
tail -3 $D/tests_multi.log; tail -c 600 $D/bench_${N}gpu.log
