# usage: bash tools/gpu_profiles.sh <tag>: the evidence set of a round - bench lines of c2-c5 and the reference arm, the launch
# list of a c2 run, ncu --set full captures of every kernel of a step on c2 / c3 / c4 / c5, the FP64 peak record
D=gpurun_out/$1
mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -2 $D/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -1 $D/smoke.log
timeout 900 python bench.py > $D/bench_c2.json 2> $D/bench_c2.err; echo "c2 rc=$?"
for c in c3 c4 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline --no-cli > $D/bench_$c.json 2> $D/bench_$c.err; echo "$c rc=$?"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err; echo "ref rc=$?"
python - > $D/fp64_peak.json <<'PY'
import ctypes as C, json, subprocess, sys
sys.path.insert(0, '.')
from xpore_b200 import _lib
L = _lib.lib()
vals = []
for _ in range(5):
    p = C.c_double(0.0)
    _lib.check(L.xpb200_measure_fp64_peak(0, 5, C.byref(p)))
    vals.append(p.value)
q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active,power.draw", "--format=csv,noheader"],
                   capture_output=True, text=True).stdout.strip()
print(json.dumps({"what": "xpb200_measure_fp64_peak: 8 independent DFMA chains per thread, 8 CTAs x 256 threads per SM, best of 5 per call",
                  "tflops": vals, "theoretical_tflops_at_1965MHz": 148 * 64 * 2 * 1.965e9 / 1e12, "nvidia_smi_after": q}))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cli > $D/ncu_launch.log 2>&1; echo "launch list rc=$?"
cap() { # config positions count
  cfg=$1; np=$2; cnt=$3
  pos=""; [ "$np" != "0" ] && pos="--positions $np"
  timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"xp_(prefilter|worklist|fit|stats)" -c $cnt -f -o $D/step_$cfg \
    python bench.py --config $cfg $pos --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_step_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
    # summarised on the box: the reports of four configurations exceed what gpurun copies back
    python tools/ncu_summary.py $D/step_$cfg.ncu-rep > $D/step_$cfg.ncu_summary.txt 2>&1
    rm -f $D/step_$cfg.ncu-rep
}
cap c2 0 9
cap c3 0 9
cap c4 0 10
cap c5 1000000 9
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xp_(pack|count|concat)" -c 4 -f -o $D/pack_c2 \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_pack_c2.log 2>&1; echo "ncu pack rc=$?"
python tools/ncu_summary.py $D/pack_c2.ncu-rep > $D/pack_c2.ncu_summary.txt 2>&1
rm -f $D/pack_c2.ncu-rep
ls -la $D
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xp_fit_tm" -c 1 -f -o $D/fit_tm_c2 \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli > $D/ncu_fit_tm_c2.log 2>&1; echo "ncu fit_tm rc=$?"
python tools/ncu_loops.py $D/fit_tm_c2.ncu-rep 2.0 > $D/fit_tm_c2.loops.txt 2>&1
du -sh $D
