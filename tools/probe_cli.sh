# usage: bash tools/probe_cli.sh <genes>: the CLI on a c2-shaped dataset, several thread counts and repeats, stage timings
TIMEFORMAT="wall %R s user %U sys %S"
set -x
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
for t in 16 16 1 16 4 16; do
  rm -rf /dev/shm/probe/out
  time python -m xpore_b200.cli diffmod --config /dev/shm/probe/config.yml --n_processes $t --timing /dev/shm/probe/t.json > /dev/null
  cat /dev/shm/probe/t.json; echo
done
rm -rf /dev/shm/probe
