"""Host ingest throughput: data.json -> CSR batch, native reader (csrc/xp_ingest.cpp) vs the Python
statement of the reference's per-gene path (json.loads + BatchBuilder.add_gene, i.e. what
xpore/scripts/diffmod.py:156-169 + xpore/diffmod/io.py:21-90 cost per gene).

    python tools/bench_ingest.py [--genes 400] [--threads 1,2,4,8]

Writes a c2-shaped dataset (2x3 runs, 100 positions per gene, 20-100 reads per run) to a temp dir.
"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import json  # noqa: E402
from collections import OrderedDict  # noqa: E402

from xpore_b200.batch import BatchBuilder, Layout  # noqa: E402
from xpore_b200.config import Configurator  # noqa: E402
from xpore_b200.diffmod import get_ids, read_index  # noqa: E402
from xpore_b200.ingest import NativeIngest  # noqa: E402
from xpore_b200.kmer_model import KmerModel  # noqa: E402
from xpore_b200.synth import config_spec, write_dataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=400)
    ap.add_argument("--threads", default="1,2,4,8")
    args = ap.parse_args()
    spec = config_spec("c2")
    spec.n_genes = args.genes
    km = KmerModel.load()
    with tempfile.TemporaryDirectory() as root:
        cfg_path = write_dataset(root, spec, km)
        cfg = Configurator(cfg_path)
        data_info, method, crit = cfg.get_data_info(), cfg.get_method(), cfg.get_criteria()
        lay = Layout.from_data_info(data_info, method["pooling"])
        f_index = OrderedDict((run, read_index(d)) for c, rd in data_info.items() for run, d in rd.items())
        ids = get_ids(f_index, data_info)
        nbytes = sum(e - s for fi in f_index.values() for s, e in fi.values())
        lo, hi = crit["readcount_min"], crit["readcount_max"]
        # Python path
        t0 = time.perf_counter()
        handles = {run: open(os.path.join(d, "data.json")) for c, rd in data_info.items() for run, d in rd.items()}
        bb = BatchBuilder(lay, km, lo, hi)
        for idx in ids:
            dd = OrderedDict()
            for c, rd in data_info.items():
                for run in rd:
                    s, e = f_index[run][idx]
                    handles[run].seek(s)
                    dd[(c, run)] = json.loads(handles[run].read(e - s))
            bb.add_gene(idx, dd)
        ref = bb.build()
        t_py = time.perf_counter() - t0
        for h in handles.values():
            h.close()
        print("dataset: %d genes, %d positions, %d reads, %.1f MB of data.json" %
              (len(ids), ref.n_positions, ref.n_reads, nbytes / 1e6))
        print("python  json.loads + BatchBuilder : %7.2f s  %8.0f positions/s  %7.1f MB/s" %
              (t_py, ref.n_positions / t_py, nbytes / 1e6 / t_py))
        ni = NativeIngest(data_info, lay, km, lo, hi)
        for nt in [int(x) for x in args.threads.split(",")]:
            t0 = time.perf_counter()
            got = ni.read(ids, f_index, n_threads=nt)
            dt = time.perf_counter() - t0
            assert got.y.tobytes() == ref.y.tobytes() and (got.grp_len == ref.grp_len).all()
            print("native  %2d thread(s)              : %7.2f s  %8.0f positions/s  %7.1f MB/s  (%.0fx)" %
                  (nt, dt, got.n_positions / dt, nbytes / 1e6 / dt, t_py / dt))
        ni.close()


if __name__ == "__main__":
    main()
