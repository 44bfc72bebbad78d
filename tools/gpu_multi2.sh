# usage: bash tools/gpu_multi2.sh <tag> <N>: the chained-pipeline test (incl. the asynchronous entry point), the 2-rank CLI test, bench.py on N GPUs under torchrun
D=gpurun_out/$1; N=$2
mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chained or two_ranks or rows_equals" > $D/tests_multi.log 2>&1; echo "rc=$?" >> $D/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N > $D/bench_${N}gpu.log 2>&1; echo "rc=$?" >> $D/bench_${N}gpu.log
tail -4 $D/tests_multi.log; grep -o '"ms_per_step": [0-9.]*' $D/bench_${N}gpu.log | tr '\n' ' '; tail -1 $D/bench_${N}gpu.log
