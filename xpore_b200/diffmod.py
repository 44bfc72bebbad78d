"""``xpore diffmod`` driver: data.json/data.index -> CSR batches -> GPU -> diffmod.table.

Mirrors ``xpore/scripts/diffmod.py:79-189`` (``diffmod(args)``) at the file level — same YAML,
same inputs, same ``diffmod.table`` / ``diffmod.log`` (header row, ``--- DIFFMOD ---``, one id per
line, ``--- SUCCESSFULLY FINISHED ---``), same ``--resume`` / ``--ids`` semantics — but replaces
the process pool of per-gene workers (``helper.py:96-116``) with "native ingest of many genes
(``xpore_b200.ingest``, C++) -> one C-ABI fit call -> single writer".  ``--n_processes`` is the
number of ingest threads that parse JSON and apply the position selection on the host; the fit
itself always runs on the GPU(s).

Multi-GPU: launch with ``torchrun``; ids are split into contiguous ranges balanced by their
``data.index`` byte spans, every rank fits its range, the fixed-width records of fitted
positions are gathered on rank 0 (NCCL ``gather``), which writes the table.
"""
from __future__ import annotations

import csv
import os
from collections import OrderedDict

import numpy as np

from .batch import Layout
from .config import Configurator
from .kmer_model import KmerModel
from .table import get_result_table_header, result_rows, write_rows


def decor_message(text):  # helper.py:44-49
    return "--- " + text.upper() + " ---\n"


def read_index(dirpath):
    """``data.index`` -> {idx: (start, end)} (diffmod.py:132-133)."""
    out = OrderedDict()
    with open(os.path.join(dirpath, "data.index")) as f:
        header = next(f, None)
        assert header is not None and header.strip().split(",")[:3] == ["idx", "start", "end"], header
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            idx, s, e = line.rsplit(",", 2)
            out[idx] = (int(s), int(e))
    return out


def get_ids(f_index, data_info):
    """ids present in at least two conditions (union over a condition's runs), sorted
    (``helper.get_ids``, helper.py:56-67)."""
    count = {}
    for cond, runs in data_info.items():
        ids = set()
        for run in runs:
            ids.update(f_index[run].keys())
        for i in ids:
            count[i] = count.get(i, 0) + 1
    return sorted(i for i, c in count.items() if c >= 2)


def diffmod(args, fit_fn=None):
    """Run the diffmod step.  ``fit_fn(batch, method) -> FitResult`` defaults to the CUDA path
    (there is no CPU implementation in this package; tests inject their checker here)."""
    from .fit import Method

    config = Configurator(args.config)
    paths = config.get_paths()
    data_info = config.get_data_info()
    method = config.get_method()
    criteria = config.get_criteria()
    km = KmerModel.load(paths["model_kmer"])
    print("Using the signal of unmodified RNA from", km.path)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    save_models = bool(getattr(args, "save_models", False))
    models_fn = None  # (batch, method) -> (FitResult, prob0, ln_prob): the fit plus the z node of every position
    if fit_fn is None:
        from .fit import fit_batch
        fit_fn = lambda b, m: fit_batch(b, m, device=local_rank)  # noqa: E731
        if save_models:
            def models_fn(b, m):
                from .fit import DeviceBatch, DeviceResult, fit_device, posterior_device
                db = DeviceBatch(b, device="cuda:%d" % local_rank)
                out = fit_device(db, m, DeviceResult(db.P, db.G, db.C, db.device, with_coef=True))
                prob0, ln_prob = posterior_device(db, out)
                return out.to_host(), prob0.cpu().numpy(), ln_prob.cpu().numpy()
    elif save_models:
        models_fn = getattr(fit_fn, "with_models", None)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                from .dist import bind_to_gpu_cpus
                bind_to_gpu_cpus(local_rank)  # pinned buffers next to this rank's GPU
            dist.init_process_group(backend)

    out_table = os.path.join(paths["out_dir"], "diffmod.table")
    out_log = os.path.join(paths["out_dir"], "diffmod.log")
    layout = Layout.from_data_info(data_info, method["pooling"])
    m = Method.from_dict(method)
    pm = m.prefilter_method

    ids_done = []
    if args.resume and os.path.exists(out_log):
        ids_done = [line.rstrip("\n") for line in open(out_log, "r")]
    elif rank == 0:
        with open(out_table, "w", newline="") as f:
            csv.writer(f, delimiter=",").writerow(get_result_table_header(layout, pm))
        with open(out_log, "w") as f:
            f.write(decor_message("diffmod"))

    f_index = OrderedDict()
    for cond, runs in data_info.items():
        for run, d in runs.items():
            f_index[run] = read_index(d)
    ids = list(args.ids) if len(args.ids) else get_ids(f_index, data_info)
    print(len(ids), "ids to be testing ...")
    done = set(ids_done)
    todo = [i for i in ids if not (args.resume and i in done)]

    if world > 1:
        from .dist import balanced_partition
        cost = np.array([sum(f_index[r][i][1] - f_index[r][i][0] for r in f_index if i in f_index[r]) for i in todo],
                        dtype=np.float64)
        mine = [todo[j] for j in balanced_partition(cost, world)[rank]]
    else:
        mine = todo

    # ---- ingest (native, --n_processes threads) -> fit -> rows, chunk by chunk ----------------------
    from .ingest import NativeIngest
    n_proc = max(1, int(args.n_processes))
    lo, hi = criteria["readcount_min"], criteria["readcount_max"]
    reader = NativeIngest(data_info, layout, km, lo, hi)
    # optional binary hand-off (xpore_b200.csrcache): built on first use, reused while the inputs are unchanged
    cache = None
    cache_path = getattr(args, "csr_cache", None)
    if cache_path:
        from .csrcache import build_cache, load_cache
        cache = load_cache(cache_path, data_info, layout, criteria, km)
        if cache is None:
            if rank == 0:
                info = build_cache(args.config, cache_path, n_threads=n_proc)
                print("csr cache written:", cache_path, info)
            if world > 1:
                dist.barrier()
            cache = load_cache(cache_path, data_info, layout, criteria, km)
    max_bytes = int(getattr(args, "batch_bytes", 0) or (256 << 20))  # data.json text per chunk
    max_genes = int(getattr(args, "batch_genes", 0) or 8192)      # and ids per chunk
    gene_rows = []  # (idx, rows) of this rank, in id order

    def flush(chunk):
        if not chunk:
            return
        if cache is not None and all(i in cache.index_of for i in chunk):
            batch = cache.batch(chunk, layout)
        else:
            batch = reader.read(chunk, f_index, n_threads=n_proc)
            batch.encode_reads()  # u16 / i32 when every read is a 2-decimal literal: 4x less PCIe traffic
        if save_models and models_fn is not None and batch.n_positions:
            from .models import save_models as write_models
            res, prob0, ln_prob = models_fn(batch, m)
            write_models(paths["models"], batch, res, prob0, ln_prob, m.dirichlet_conc)  # diffmod.py:61-63
        else:
            res = fit_fn(batch, m) if batch.n_positions else None
        rows = result_rows(batch, res, pm) if res is not None else []
        by_gene = OrderedDict((idx, []) for idx in chunk)
        for r in rows:
            by_gene[r[0]].append(r)
        if world == 1:
            for idx, rws in by_gene.items():  # rows first, then the id (keeps --resume valid)
                if rws:
                    write_rows(out_table, rws)
                with open(out_log, "a") as f:
                    f.write(idx + "\n")
        else:
            gene_rows.extend(by_gene.items())

    chunk, nbytes = [], 0
    for idx in mine:
        chunk.append(idx)
        nbytes += sum(f_index[r][idx][1] - f_index[r][idx][0] for r in f_index if idx in f_index[r])
        if nbytes >= max_bytes or len(chunk) >= max_genes:
            flush(chunk)
            chunk, nbytes = [], 0
    flush(chunk)
    reader.close()

    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(gene_rows, gathered, dst=0)
        if rank == 0:
            for per_rank in gathered:
                for idx, rws in per_rank:
                    if rws:
                        write_rows(out_table, rws)
                    with open(out_log, "a") as f:
                        f.write(idx + "\n")
        dist.barrier()
    if rank == 0:
        with open(out_log, "a+") as f:
            f.write(decor_message("successfully finished"))
