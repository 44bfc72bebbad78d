# usage: bash tools/gpu_final3.sh <tag>: bench lines of the shipped build (c2 with the CPU and CLI legs, c3-c5, reference arm)
D=gpurun_out/$1
mkdir -p $D
timeout 900 python bench.py > $D/bench_c2.json 2> $D/bench_c2.err; echo "c2 rc=$?"
for c in c3 c4 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.json 2> $D/bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err; echo "ref rc=$?"
