# usage: bash tools/gpu_pod.sh <tag>: GPU tests with the shipped library, then the experiment build with XPB200_POD = 0 / 1 / 2 / 4
D=gpurun_out/$1
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
tail -3 $D/tests.log
export XPB200_LIB=$PWD/build/exp/libxp_exp.so
for c in c2 c3 c4; do
  for pod in 0 1 2 4; do
    XPB200_POD=$pod timeout 300 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_${c}_pod$pod.log 2>&1
    echo "$c pod=$pod rc=$? $(grep -o '"ms_per_step": [0-9.]*' $D/bench_${c}_pod$pod.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_${c}_pod$pod.log) $(grep -o '"fit": [0-9.]*' $D/bench_${c}_pod$pod.log)"
  done
done
XPB200_POD=0 timeout 400 python bench.py --config c5 --positions 400000 --no-cpu-baseline --no-cli > $D/bench_c5_pod0.log 2>&1
XPB200_POD=1 timeout 400 python bench.py --config c5 --positions 400000 --no-cpu-baseline --no-cli > $D/bench_c5_pod1.log 2>&1
XPB200_POD=2 timeout 400 python bench.py --config c5 --positions 400000 --no-cpu-baseline --no-cli > $D/bench_c5_pod2.log 2>&1
for pod in 0 1 2; do echo "c5 pod=$pod $(grep -o '"ms_per_step": [0-9.]*' $D/bench_c5_pod$pod.log | head -1) $(grep -o '"sweeps_mean": [0-9.]*' $D/bench_c5_pod$pod.log) $(grep -o '"fit": [0-9.]*' $D/bench_c5_pod$pod.log)"; done
