"""``xpore diffmod`` driver: data.json/data.index -> CSR batches -> GPU -> diffmod.table.

Mirrors ``xpore/scripts/diffmod.py:79-189`` (``diffmod(args)``) at the file level — same YAML,
same inputs, same ``diffmod.table`` / ``diffmod.log`` (header row, ``--- DIFFMOD ---``, one id per
line, ``--- SUCCESSFULLY FINISHED ---``), same ``--resume`` / ``--ids`` semantics — but replaces
the process pool of per-gene workers (``helper.py:96-116``) with a three-stage pipeline over chunks
of ids:

    ingest thread   native reader (``xpore_b200.ingest``, C++, ``--n_processes`` threads): JSON text of chunk k+1
                    -> CSR arrays in page-locked staging buffers, reads as the u16 codes found in the text
    main thread     ``xpb200_fit_host_rows`` on chunk k: staged arrays -> GPU -> the records of the table's rows
    writer thread   rows of chunk k-1 -> ``diffmod.table``, then the chunk's ids -> ``diffmod.log``

so the text parsing (the end-to-end bottleneck once the fit runs on a GPU) overlaps the
GPU call and the formatting of the rows.  The fit itself always runs on the GPU(s).

Multi-GPU: launch with ``torchrun``; ids are split into contiguous ranges balanced by their
``data.index`` byte spans, every rank runs the pipeline on its range, and the fixed-width numeric records of
its rows - gene ordinal, position, k-mer row, then the record of ``xpb200_fit_host_rows`` - are gathered
on rank 0 (``dist.gather_records``: NCCL ``gather``), which writes the table.
"""
from __future__ import annotations

import csv
import os
import queue
import threading
import time
from collections import OrderedDict

import numpy as np

from .batch import Layout
from .config import Configurator
from .kmer_model import KmerModel
from .table import get_result_table_header, rows_from_records, write_rows_to


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


def records_from_dense(res, index_offset=0):
    """The records ``xpb200_fit_host_rows`` returns, cut from dense result arrays (injected checkers, --save_models)."""
    keep = np.nonzero(res.keep_row)[0]
    cols = [(keep + index_offset).astype(np.float64)[:, None], res.status[keep].astype(np.float64)[:, None],
            res.n_iter[keep].astype(np.float64)[:, None], res.pval[keep][:, None], res.fit[keep], res.stats[keep]]
    return np.concatenate(cols, axis=1) if len(keep) else np.zeros((0, 4 + res.fit.shape[1] + res.stats.shape[1]))


class _Stage(threading.Thread):
    """A pipeline stage: ``fn(item)`` for every item of ``inq`` until ``None``; the first exception is kept."""

    def __init__(self, fn, inq, name):
        super().__init__(name=name, daemon=True)
        self.fn, self.inq, self.error = fn, inq, None

    def run(self):
        while True:
            item = self.inq.get()
            if item is None:
                return
            if self.error is None:
                try:
                    self.fn(item)
                except BaseException as e:  # noqa: BLE001  (re-raised in the main thread)
                    self.error = e


def diffmod(args, fit_fn=None):
    """Run the diffmod step.  ``fit_fn(batch, method) -> FitResult`` defaults to the CUDA path
    (there is no CPU implementation in this package; tests inject their checker here).  Returns the stage timings."""
    from .fit import Method, fit_width, stats_width

    t_start = time.perf_counter()
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
    injected = fit_fn is not None
    models_fn = None  # (batch, method) -> (FitResult, prob0, ln_prob): the fit plus the z node of every position
    if not injected and save_models:
        def models_fn(b, m):
            from .fit import DeviceBatch, DeviceResult, fit_device, posterior_device
            db = DeviceBatch(b, device="cuda:%d" % local_rank)
            out = fit_device(db, m, DeviceResult(db.P, db.G, db.C, db.device, with_coef=True))
            prob0, ln_prob = posterior_device(db, out)
            return out.to_host(), prob0.cpu().numpy(), ln_prob.cpu().numpy()
    elif injected and save_models:
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
    G, Cn = layout.n_groups, layout.n_conditions
    RW = 4 + fit_width(G) + stats_width(G, Cn)

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

    def span_bytes(idx):
        return sum(f_index[r][idx][1] - f_index[r][idx][0] for r in f_index if idx in f_index[r])

    if world > 1:
        from .dist import balanced_partition
        cost = np.array([span_bytes(i) for i in todo], dtype=np.float64)
        parts = balanced_partition(cost, world)
        mine = [todo[j] for j in parts[rank]]
    else:
        parts = None
        mine = todo

    from .ingest import NativeIngest, Staging
    n_proc = max(1, int(args.n_processes))
    lo, hi = criteria["readcount_min"], criteria["readcount_max"]
    reader = NativeIngest(data_info, layout, km, lo, hi)
    # optional binary hand-off (xpore_b200.csrcache): built on first use, reused while the inputs are unchanged
    cache = None
    cache_path = getattr(args, "csr_cache", None)
    if cache_path:
        from .csrcache import build_cache, load_cache
        cache = load_cache(cache_path, data_info, layout, criteria, km)
        # the decision to build is collective: every rank learns whether rank 0 found a usable cache, so that either
        # all of them meet in the barrier below or none does (a rank that arrives after rank 0 has written the file
        # would otherwise skip a barrier rank 0 is waiting in)
        build = cache is None
        if world > 1:
            flag = [build]
            dist.broadcast_object_list(flag, src=0)
            build = bool(flag[0])
        if build:
            if rank == 0:
                info = build_cache(args.config, cache_path, n_threads=n_proc)
                print("csr cache written:", cache_path, info)
            if world > 1:
                dist.barrier()
            cache = load_cache(cache_path, data_info, layout, criteria, km)
    max_bytes = int(getattr(args, "batch_bytes", 0) or (256 << 20))  # data.json text per chunk
    max_genes = int(getattr(args, "batch_genes", 0) or 8192)      # and ids per chunk
    staged = not injected and models_fn is None  # u16 codes in page-locked buffers, no float64 copy of the reads
    timing = {"ingest_s": 0.0, "gpu_s": 0.0, "rows_s": 0.0, "write_s": 0.0, "text_bytes": 0, "positions": 0, "reads": 0,
              "fitted": 0, "rows": 0, "chunks": 0, "near_threshold": 0}
    gathered_rows = []  # world > 1: numeric records of this rank's rows, gathered at the end

    # ---- chunks of ids --------------------------------------------------------------------------------
    chunks, chunk, nbytes = [], [], 0
    for idx in mine:
        chunk.append(idx)
        nbytes += span_bytes(idx)
        if nbytes >= max_bytes or len(chunk) >= max_genes:
            chunks.append((chunk, nbytes))
            chunk, nbytes = [], 0
    if chunk:
        chunks.append((chunk, nbytes))
    ordinal = {idx: i for i, idx in enumerate(ids)}
    kmer_row_of = {k: i for i, k in enumerate(km.kmers)} if world > 1 else None

    # ---- stage 3: rows -> diffmod.table, ids -> diffmod.log (or: keep the numeric records for the gather) ----
    def write_stage(item):
        chunk_ids, keys, records, present, krow = item
        t0 = time.perf_counter()
        if world > 1:
            if len(records):
                pidx = records[:, 0].astype(np.int64)
                kk = [keys.key(int(p)) for p in pidx]
                pos = np.asarray([_encode_position(k[1], odd_positions) for k in kk], dtype=np.float64)
                gene = np.asarray([ordinal[k[0]] for k in kk], dtype=np.float64)
                head = np.stack([gene, pos, krow, _pack_mask(present)], axis=1)
                gathered_rows.append(np.concatenate([head, records], axis=1))
            timing["rows_s"] += time.perf_counter() - t0
            return
        pidx = records[:, 0].astype(np.int64) if len(records) else np.zeros(0, np.int64)
        kk = [keys.key(int(p)) for p in pidx]
        rows = rows_from_records(records, kk, present, layout, pm)
        t1 = time.perf_counter()
        by_gene = OrderedDict((idx, []) for idx in chunk_ids)
        for r in rows:
            by_gene[r[0]].append(r)
        with open(out_table, "a", newline="") as ft, open(out_log, "a") as fl:
            for idx, rws in by_gene.items():  # rows first, then the id (keeps --resume valid)
                if rws:
                    write_rows_to(ft, rws)
                    ft.flush()
                fl.write(idx + "\n")
        timing["rows_s"] += t1 - t0
        timing["write_s"] += time.perf_counter() - t1

    odd_positions = []  # position strings that are not canonical integers (never with dataprep output)
    write_q = queue.Queue(maxsize=4)
    writer = _Stage(write_stage, write_q, "xpb200-writer")
    writer.start()

    # ---- stage 1: ingest --------------------------------------------------------------------------------
    n_sets = 2
    free_q = queue.Queue()
    for _ in range(n_sets):
        free_q.put(Staging() if staged else None)
    work_q = queue.Queue(maxsize=n_sets)

    ingest_state = {"device_set": False}

    def ingest_stage(item):
        chunk_ids, nb = item
        st = free_q.get()
        t0 = time.perf_counter()

        def before_alloc():  # after the parse, before the first page-locked allocation of this thread
            if staged and not ingest_state["device_set"]:
                from . import _lib
                t1 = time.perf_counter()
                timing["cuda_init_s"] = _lib.warm_up_wait()  # the context exists (created under the imports and the parse)
                timing["cuda_init_wait_s"] = time.perf_counter() - t1
                _lib.lib().xpb200_set_device(local_rank)  # this thread's allocations belong to this rank's GPU
                ingest_state["device_set"] = True
        if cache is not None and all(i in cache.index_of for i in chunk_ids):
            batch = cache.batch(chunk_ids, layout)
        else:
            batch = reader.read(chunk_ids, f_index, n_threads=n_proc, staging=st, before_alloc=before_alloc)
            timing["ingest_parse_s"] = timing.get("ingest_parse_s", 0.0) + reader.last_times["parse_s"]
            timing["ingest_copy_s"] = timing.get("ingest_copy_s", 0.0) + reader.last_times["copy_s"]
            if batch.y is not None:
                batch.encode_reads()  # u16 / i32 when every read is a 2-decimal literal: 4x less PCIe traffic
        timing["ingest_s"] += time.perf_counter() - t0
        timing["text_bytes"] += nb
        work_q.put((chunk_ids, batch, st))

    if staged and chunks:
        # the CUDA context (0.3 - 1.5 s to create) comes up on a background thread - started at process entry by the CLI,
        # here otherwise - while the ingest thread parses the first chunk
        from . import _lib
        _lib.warm_up_async(local_rank)

    ingest_q = queue.Queue()
    ingester = _Stage(ingest_stage, ingest_q, "xpb200-ingest")
    ingester.start()
    for c in chunks:
        ingest_q.put(c)
    ingest_q.put(None)
    # ---- stage 2: the GPU call (this thread) -----------------------------------------------------------
    def gpu_stage(chunk_ids, batch, st):
        from .fit import fit_rows
        t0 = time.perf_counter()
        P = batch.n_positions
        n_fitted = 0
        if P == 0:
            records = np.zeros((0, RW))
        elif models_fn is not None:
            from .models import save_models as write_models
            res, prob0, ln_prob = models_fn(batch, m)
            write_models(paths["models"], batch, res, prob0, ln_prob, m.dirichlet_conc)  # diffmod.py:61-63
            records, n_fitted = records_from_dense(res), res.n_fitted
        elif injected:
            res = fit_fn(batch, m)
            records, n_fitted = records_from_dense(res), res.n_fitted
        else:
            rows, n_fitted = fit_rows(batch, m, device=local_rank, staging=st)
            records = rows.copy()  # the staging buffers go back to the ingest thread
        if len(records):
            near = (records[:, 1].astype(np.int64) & 512) != 0
            timing["near_threshold"] += int(near.sum())
        pidx = records[:, 0].astype(np.int64) if len(records) else np.zeros(0, np.int64)
        present = np.asarray(batch.grp_len)[pidx] > 0
        keys = batch.keys if batch.keys is not None else _ListKeys(batch)
        krow = None
        if world > 1:  # row of every record's k-mer in the model: the numeric stand-in for the string in the gather
            krow = batch.kmer_row[pidx].astype(np.float64) if batch.kmer_row is not None else \
                np.asarray([kmer_row_of[batch.kmers[int(p)]] for p in pidx], dtype=np.float64)
        timing["gpu_s"] += time.perf_counter() - t0
        if timing["chunks"] == 0:
            timing["gpu_first_chunk_s"] = timing["gpu_s"]  # includes device buffers, streams, kernel module load
        timing["positions"] += P
        timing["reads"] += batch.n_reads
        timing["fitted"] += int(n_fitted)
        timing["rows"] += len(records)
        timing["chunks"] += 1
        free_q.put(st)
        write_q.put((chunk_ids, keys, records, present, krow))

    try:
        for _ in range(len(chunks)):
            while True:
                try:
                    item = work_q.get(timeout=0.2)
                    break
                except queue.Empty:
                    if ingester.error is not None:
                        raise ingester.error
            if writer.error is not None:
                raise writer.error
            gpu_stage(*item)
    finally:
        write_q.put(None)
        writer.join()
        ingester.join(timeout=1.0)
        reader.close()
    for stage in (ingester, writer):
        if stage.error is not None:
            raise stage.error

    if world > 1:
        import torch

        from .dist import gather_records
        W = 4 + RW
        mine_rec = np.concatenate(gathered_rows, axis=0) if gathered_rows else np.zeros((0, W))
        t = torch.from_numpy(np.ascontiguousarray(mine_rec))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        allrec = gather_records(t, rank, world)  # the data plane: fixed-width f64 records, rank order
        odd_all = [None] * world if rank == 0 else None
        dist.gather_object(odd_positions, odd_all, dst=0)  # normally empty lists
        if rank == 0:
            t0 = time.perf_counter()
            allrec = allrec.cpu().numpy()
            counts = gather_records.last_counts
            with open(out_table, "a", newline="") as ft, open(out_log, "a") as fl:
                o = 0
                for r in range(world):
                    rec = allrec[o:o + counts[r]]
                    o += counts[r]
                    kk = [(ids[int(g)], _decode_position(p, odd_all[r]), km.kmers[int(k)]) for g, p, k in rec[:, :3]]
                    rows = rows_from_records(rec[:, 4:], kk, _unpack_mask(rec[:, 3], G), layout, pm)
                    by_gene = OrderedDict((todo[j], []) for j in parts[r])
                    for row in rows:
                        by_gene[row[0]].append(row)
                    for idx, rws in by_gene.items():
                        if rws:
                            write_rows_to(ft, rws)
                        fl.write(idx + "\n")
            timing["write_s"] += time.perf_counter() - t0
        dist.barrier()
    if rank == 0:
        with open(out_log, "a+") as f:
            f.write(decor_message("successfully finished"))
    timing["total_s"] = time.perf_counter() - t_start
    timing["ingest_threads"] = n_proc
    if timing["near_threshold"]:
        print("note: %d emitted position(s) had a prefilter p-value within 1e-9 (relative) of the threshold"
              % timing["near_threshold"])
    tpath = getattr(args, "timing", None)
    if tpath and rank == 0:
        import json
        with open(tpath, "w") as f:
            json.dump(timing, f)
    return timing


class _ListKeys:
    """Keys of a batch that carries plain lists (cache, builders)."""

    def __init__(self, batch):
        self.b = batch

    def key(self, p):
        return (self.b.ids[p], self.b.positions[p], self.b.kmers[p])


def _encode_position(s, odd):
    """Position string -> float64: the integer itself when the string is its canonical form (what dataprep writes),
    else -(1 + index) into the rank's list of odd strings."""
    if s.isdigit() and (s == "0" or s[0] != "0") and len(s) <= 15:
        return float(int(s))
    odd.append(s)
    return -float(len(odd))


def _decode_position(v, odd):
    return str(int(v)) if v >= 0 else odd[int(-v) - 1]


def _pack_mask(present):
    """[n, G <= 64] bool -> one float64 per row (exact below 2^53: G <= 52 here; larger G uses two halves)."""
    present = np.asarray(present, dtype=bool)
    n, G = present.shape if present.ndim == 2 else (0, 0)
    if n == 0:
        return np.zeros(0)
    assert G <= 52, "more than 52 group slots: gather the rows with --n_gpus 1"
    w = (1 << np.arange(G, dtype=np.int64))[None, :]
    return (present * w).sum(axis=1).astype(np.float64)


def _unpack_mask(v, G):
    bits = v.astype(np.int64)[:, None] >> np.arange(G, dtype=np.int64)[None, :]
    return (bits & 1).astype(bool)
