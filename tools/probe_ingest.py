"""Where the CLI's ingest time goes on a given box: native reader per thread count (pageable and page-locked
staging), the cost of the first page-locked allocation (CUDA context), and malloc tunables.
    python tools/probe_ingest.py [genes]"""
import os
import sys
import time
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from xpore_b200.batch import Layout  # noqa: E402
from xpore_b200.config import Configurator  # noqa: E402
from xpore_b200.diffmod import get_ids, read_index  # noqa: E402
from xpore_b200.ingest import NativeIngest, Staging  # noqa: E402
from xpore_b200.kmer_model import KmerModel  # noqa: E402
from xpore_b200.synth import config_spec  # noqa: E402

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
spec = config_spec("c2")
km = KmerModel.load()
root = bench.scratch_dir("xpb200_probe_")
cfg_path, batch, nbytes = bench.write_files(spec, km, genes, root)
cfg = Configurator(cfg_path)
data_info, method, crit = cfg.get_data_info(), cfg.get_method(), cfg.get_criteria()
lay = Layout.from_data_info(data_info, method["pooling"])
t0 = time.perf_counter()
f_index = OrderedDict((run, read_index(d)) for c, rd in data_info.items() for run, d in rd.items())
ids = get_ids(f_index, data_info)
print("index %.3f s; %d ids, %.1f MB text, cores %d" % (time.perf_counter() - t0, len(ids), nbytes / 1e6, len(os.sched_getaffinity(0))))
ni = NativeIngest(data_info, lay, km, crit["readcount_min"], crit["readcount_max"])
for nt in (1, 2, 4, 8, 16, 32):
    t0 = time.perf_counter()
    b = ni.read(ids, f_index, n_threads=nt)
    dt = time.perf_counter() - t0
    print("pageable  threads %2d: %.3f s  %.0f MB/s" % (nt, dt, nbytes / 1e6 / dt), flush=True)
st = Staging()
t0 = time.perf_counter()
st.array("y", "u2", (1024,))
print("first page-locked allocation (CUDA context): %.3f s" % (time.perf_counter() - t0))
for nt in (1, 4, 16):
    for rep in range(2):
        t0 = time.perf_counter()
        b = ni.read(ids, f_index, n_threads=nt, staging=st)
        dt = time.perf_counter() - t0
        print("staged    threads %2d (pass %d): %.3f s  %.0f MB/s" % (nt, rep, dt, nbytes / 1e6 / dt), flush=True)
import shutil
shutil.rmtree(root, ignore_errors=True)
