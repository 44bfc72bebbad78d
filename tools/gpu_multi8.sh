# usage: bash tools/gpu_multi8.sh <tag> <N>: bench.py on N GPUs under torchrun
D=gpurun_out/$1; N=$2
mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N > $D/bench_${N}gpu.log 2>&1; echo "rc=$?" >> $D/bench_${N}gpu.log
grep -o '"ms_per_step": [0-9.]*' $D/bench_${N}gpu.log | tr '\n' ' '; tail -1 $D/bench_${N}gpu.log
