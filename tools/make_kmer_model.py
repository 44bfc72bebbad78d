"""Convert the unmodified-RNA k-mer signal table that xPore ships into the
binary form this package loads (``xpore_b200/data/model_kmer.npz``).

The table is *input data* for the diffmod path (priors and ``conf_mu``; reference
``xpore/diffmod/configurator.py:19-22`` picks it when the YAML has no ``prior:`` key,
``xpore/scripts/diffmod.py:23-48`` turns each row into Normal-Gamma priors).
It is not source code: 1025 rows of (5-mer, mean, stdv).  Run in the build
container only (needs /root/reference); the .npz is what travels.

    python tools/make_kmer_model.py [/root/reference/xpore/diffmod/model_kmer.csv]
"""
import csv
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/xpore/diffmod/model_kmer.csv"
kmers, mean, stdv = [], [], []
with open(src) as f:
    rd = csv.reader(f)
    header = next(rd)
    assert header == ["model_kmer", "model_mean", "model_stdv"], header
    for k, m, s in rd:
        kmers.append(k)
        mean.append(float(m))
        stdv.append(float(s))
out = os.path.join(os.path.dirname(__file__), "..", "xpore_b200", "data", "model_kmer.npz")
np.savez_compressed(out, kmer=np.array(kmers, dtype="S5"), mean=np.array(mean), stdv=np.array(stdv))
print("wrote", os.path.abspath(out), len(kmers), "rows")
