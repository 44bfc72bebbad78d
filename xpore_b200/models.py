"""``xpore diffmod --save_models``: per-id dumps of the fitted models.

The reference writes ``<out>/models/<idx>.hdf5`` with one group per position holding the attribute ``kmer``
and ``nodes/<node>/<field>`` datasets (``io.save_models_to_hdf5``, ``xpore/diffmod/io.py:99-139``)::

    <idx>/<position>                       attrs['kmer']
    <idx>/<position>/nodes/x               data  [n, G] bool   group_names [G]
    <idx>/<position>/nodes/y               data  [n]           N [G, 2]  mean [2]  variance [2]
    <idx>/<position>/nodes/z               prob  [n, 2]        ln_prob [n, 2]
    <idx>/<position>/nodes/w               concentration [G, 2]
    <idx>/<position>/nodes/mu_tau          location lambda alpha beta  [2] each
    <idx>/<position>/nodes/y_run_names     data  [n]
    <idx>/<position>/nodes/y_condition_names  data [n]

:func:`position_nodes` builds exactly these fields from a fitted batch: the Normal-Gamma and Dirichlet
parameters and ``N`` come from the fit records, the z node from ``xpb200_posterior_device`` (responsibilities
of the last E-step, recomputed on the GPU from the coefficients the fit kernel keeps when asked to), and the
y-node mean / variance from those responsibilities in the reference's own formula (``gmm.py:179-189``).
:func:`save_models` writes them as ``<idx>.hdf5`` when ``h5py`` is importable and always as ``<idx>.npz``
(flat keys ``<position>/nodes/<node>/<field>`` and ``<position>/kmer``) - h5py is not part of this image, so
the ``.npz`` is what the tests read back.
"""
from __future__ import annotations

import os

import numpy as np

from .batch import Batch
from .fit import FIT_ALPHA, FIT_BETA, FIT_LAM, FIT_MU, FIT_N, FitResult


def position_nodes(batch: Batch, res: FitResult, p: int, prob0: np.ndarray, ln_prob: np.ndarray, conc0: float):
    """The reference's node dictionary of fitted position ``p`` (only the groups present there)."""
    lay = batch.layout
    a, b = int(batch.read_off[p]), int(batch.read_off[p + 1])
    y = np.asarray(batch.y[a:b], dtype=np.float64)
    glen = batch.grp_len[p]
    # the reference's model groups are the columns of get_dummies: present groups in SORTED label order
    # (io.py:12-18), while the reads stay run by run in config order (io.py:46-58) - as they are in the batch
    present = np.asarray(sorted(np.nonzero(glen > 0)[0], key=lambda g: lay.group_names[g]), dtype=np.int64)
    group_of_read = np.repeat(np.arange(len(glen)), glen)
    col = {g: j for j, g in enumerate(present)}
    x = np.zeros((len(y), len(present)), dtype=bool)
    x[np.arange(len(y)), [col[g] for g in group_of_read]] = True
    r0 = np.asarray(prob0[a:b], dtype=np.float64)
    prob = np.stack([r0, 1 - r0], axis=-1)  # gmm.py:242
    N = res.fit[p, FIT_N:].reshape(-1, 2)[present]
    Nk = N.sum(axis=-2)
    Nk[Nk < 1e-30] = 0.0  # gmm.py:181
    Ninv = np.divide(1.0, Nk, out=np.zeros_like(Nk), where=Nk != 0.0)
    mean = Ninv * np.sum(prob * y[:, None], axis=-2)
    var = Ninv * np.sum(prob * (y[:, None] - mean[None, :]) ** 2, axis=-2)
    run_names = np.asarray(lay.run_names if not lay.pooling else [])
    if lay.pooling:
        # Deviation from the reference's dump: pooled batches keep one slot per condition and their reads condition by
        # condition; which run a read came from is not part of the batch, so `y_run_names` is not written, and the
        # reads of a position are ordered by condition (the reference: run by run in config order, io.py:46-58).
        y_cond = np.asarray(lay.group_names)[group_of_read]
        y_run = None
    else:
        y_run = run_names[group_of_read]
        y_cond = np.asarray(lay.condition_names)[np.asarray(lay.grp_cond)[group_of_read]]
    nodes = {
        "x": {"data": x, "group_names": np.asarray(lay.group_names)[present]},
        "y": {"data": y, "N": N, "mean": mean, "variance": var},
        "z": {"prob": prob, "ln_prob": np.asarray(ln_prob[a:b], dtype=np.float64)},
        "w": {"concentration": conc0 + N},
        "mu_tau": {"location": res.fit[p, FIT_MU:FIT_MU + 2], "lambda": res.fit[p, FIT_LAM:FIT_LAM + 2],
                   "alpha": res.fit[p, FIT_ALPHA:FIT_ALPHA + 2], "beta": res.fit[p, FIT_BETA:FIT_BETA + 2]},
        "y_condition_names": {"data": y_cond},
    }
    if y_run is not None:
        nodes["y_run_names"] = {"data": y_run}
    return nodes


def save_models(models_dir: str, batch: Batch, res: FitResult, prob0, ln_prob, conc0: float = 0.001):
    """Write one file per id of ``batch`` with every fitted position (``diffmod.py:61-63``).  Returns the paths."""
    os.makedirs(models_dir, exist_ok=True)
    try:
        import h5py  # noqa: F401
    except Exception:
        h5py = None
    by_id = {}
    for p in np.nonzero(res.selected)[0]:
        if res.n_iter[p] < 1:
            continue
        by_id.setdefault(batch.ids[p], []).append(int(p))
    paths = []
    for idx, plist in by_id.items():
        flat = {}
        for p in plist:
            pos, kmer = str(batch.positions[p]), batch.kmers[p]
            flat["%s/kmer" % pos] = np.asarray(kmer)
            for node, fields in position_nodes(batch, res, p, prob0, ln_prob, conc0).items():
                for name, value in fields.items():
                    flat["%s/nodes/%s/%s" % (pos, node, name)] = np.asarray(value)
        path = os.path.join(models_dir, "%s.npz" % idx)
        np.savez(path, **flat)
        paths.append(path)
        if h5py is not None:  # the reference's own container and layout
            with h5py.File(os.path.join(models_dir, "%s.hdf5" % idx), "w") as f:
                g = f.create_group(idx)
                for key, value in flat.items():
                    pos, rest = key.split("/", 1)
                    gp = g.require_group(pos)
                    if rest == "kmer":
                        gp.attrs["kmer"] = str(value).encode("UTF-8")
                        continue
                    if value.dtype.kind == "U":
                        value = np.asarray([v.encode("UTF-8") for v in value.ravel()]).reshape(value.shape)
                    gp[rest] = value
    return paths
