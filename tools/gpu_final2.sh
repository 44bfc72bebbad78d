# usage: bash tools/gpu_final2.sh <tag>: bench lines of the shipped build (c2 with the CPU and CLI legs, c3-c5), reference arm, timeline of the chained pipeline
D=gpurun_out/$1
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -2 $D/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -1 $D/smoke.log
timeout 900 python bench.py > $D/bench_c2.json 2> $D/bench_c2.err; echo "c2 rc=$?"
for c in c3 c4 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.json 2> $D/bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err; echo "ref rc=$?"
XPB200_LIB=$PWD/build/exp/libxp_exp.so XPB200_CHAIN_POST=0 XPB200_TRACE=1 timeout 400 python bench.py --config c2 --no-cpu-baseline --no-cli > $D/bench_c2_trace.log 2> $D/trace_c2.log
tail -29 $D/trace_c2.log > $D/chain_timeline_c2.txt
