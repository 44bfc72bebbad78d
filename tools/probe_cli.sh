# usage: bash tools/probe_cli.sh <genes>: the CLI on a c2-shaped dataset, several thread counts and repeats, stage timings;
# then the same while another process keeps a CUDA context open (what persistence mode / a busy box looks like)
TIMEFORMAT="wall %R s user %U sys %S"
nvidia-smi -q | grep -i -m2 "persistence"
python - <<PY
import sys, os
sys.path.insert(0, '.')
import bench
from xpore_b200.kmer_model import KmerModel
from xpore_b200.synth import config_spec
os.makedirs('/dev/shm/probe', exist_ok=True)
cfg, b, n = bench.write_files(config_spec('c2'), KmerModel.load(), $1, '/dev/shm/probe')
print(cfg, b.n_positions, n)
PY
run() {
  for t in "$@"; do
    rm -rf /dev/shm/probe/out
    time python -m xpore_b200.cli diffmod --config /dev/shm/probe/config.yml --n_processes $t --timing /dev/shm/probe/t.json > /dev/null
    python -c "
import json; t=json.load(open('/dev/shm/probe/t.json'))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in t.items() if k.endswith('_s') or k=='ingest_threads'})"
  done
}
echo "== no other context"; run 16 8 16 8 16
rm -rf /dev/shm/probe
