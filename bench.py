#!/usr/bin/env python
"""Benchmark of the diffmod hot path: positions fitted per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--positions P]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the unmodified reference CLI on all host cores

One step = one pass of the whole path (prefilter t-test -> worklist -> VB-EM fit -> statistics)
over one synthetic batch of the BASELINE.json shape (default c2: 2 conditions x 3 replicates,
1M positions, 20-100 reads per run, t-test prefilter at 0.1).

`value`  positions fitted/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    the same metric through the C ABI with HOST buffers: pinned-host -> device copies of the batch and
         device -> host copies of the records of the table's rows inside the timed region.
`roofline`  the fit kernel against the measured FP64 DFMA peak (this path is FP64-pipe bound,
         SURVEY.md 8d), plus its HBM figures.
`cpu_baseline`  the reference's CPU path on all host cores, on a bounded sample of the same workload: the
         unmodified reference (oracle/_ref, kind "reference") when that copy exists, else the numpy/scipy oracle port.
`cli_e2e`  (N = 1) the drop-in CLI, wall clock: real data.json / data.index files -> `python -m xpore_b200.cli diffmod`
         -> diffmod.table, with the stage breakdown, next to the reference CLI on a subset of the same files.
Multi-GPU: each rank fits its own shard (weak scaling); the rows of every rank are gathered on rank 0 over NCCL
inside the timed step.  `config.scaling_detail` adds the strong-scaling measurement: the same 1M-position batch cut by
the coverage-balanced partition across the ranks.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

FLOP_PER_READ_ITER = 100.0   # 35 plain flop + exp(25) + log1p(30) + div(10)
FLOP_PER_SPECIAL = 60.0      # digamma / lgamma, (6G+4) of them per position-sweep
FP64_THEORETICAL = lambda sms, mhz: sms * 64 * 2 * mhz * 1e6 / 1e12  # noqa: E731  64 DFMA lanes per SM and clock

WORKLOADS = {"c1": "c1: 1 gene x 100 positions, 2x2, 50-200 reads/run, t-test prefilter 0.1",
             "c2": "c2: HEK293T WT vs METTL3-KO shape, 2 conditions x 3 replicates, 1M positions, "
                   "20-100 reads/run, synthetic reads, t-test prefiltering 0.1",
             "c3": "c3: c2 shape, pooling=True, no prefilter, max_iters=500",
             "c4": "c4: 100k positions, 2x2, 20-5000 reads/position log-uniform",
             "c5": "c5: 6 conditions x 3 replicates, 2M positions, 20-60 reads/run"}


# ------------------------------------------------------------------------------------------
# CPU arm, port flavour: the oracle, one task = a chunk of positions (the reference's task = one gene)
# ------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    """Prefilter + fit + statistics row for a chunk of positions with the oracle."""
    from oracle import xpore_oracle as orc
    ys, glens, kms, kss, gc, names, data_info, pooling, pre, max_iters, stop = args
    rng = np.random.RandomState(0)
    fitted = 0
    for y, gl, km, ks in zip(ys, glens, kms, kss):
        present = np.nonzero(gl > 0)[0]
        pval = None
        if pre:
            x, _ = orc.one_hot(np.repeat(gc, gl))
            pval = orc.prefilter_pvalue(y, x, pre["method"])
            if not (np.isnan(pval) or pval < pre["threshold"]):
                continue
        X = np.zeros((len(y), len(present)), dtype=bool)
        o = 0
        for j, g in enumerate(present):
            X[o:o + gl[g], j] = True
            o += gl[g]
        f = orc.fit_position(y, X, km, ks, max_iters, stop, rng=rng)
        orc.result_row(("id", "0", "NNNNN"), f, [names[g] for g in present], km, ks, data_info, pooling, pval)
        fitted += 1
    return fitted


def cpu_baseline_port(batch, spec, n_positions, cores=None, chunk=25):
    """Time the oracle on the first n_positions of the batch with a process pool over all host
    cores; returns dict(value=fitted positions/s, ...)."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    n_positions = min(n_positions, batch.n_positions)
    lay = batch.layout
    data_info = spec.data_info()
    pre = spec.method.get("prefiltering", False)
    tasks = []
    for c0 in range(0, n_positions, chunk):
        idx = range(c0, min(c0 + chunk, n_positions))
        tasks.append(([batch.y[batch.read_off[p]:batch.read_off[p + 1]].copy() for p in idx],
                      [batch.grp_len[p].copy() for p in idx], [float(batch.kmer_mean[p]) for p in idx],
                      [float(batch.kmer_std[p]) for p in idx], lay.grp_cond, lay.group_names, data_info,
                      lay.pooling, pre, int(spec.method.get("max_iters", 500)),
                      float(spec.method.get("stopping_criteria", 1e-5))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, tasks[:cores])  # warm the workers (imports)
        t0 = time.perf_counter()
        fitted = sum(pool.map(_cpu_chunk, tasks, chunksize=1))
        dt = time.perf_counter() - t0
    return {"value": fitted / dt, "unit": "positions fitted/s", "cores": cores, "kind": "port",
            "sample": "first %d positions of the same batch (%d fitted) through oracle/xpore_oracle.py "
                      "(numpy/scipy restatement, same per-position call structure as the reference) "
                      "with a %d-process pool, %.1f s wall" % (n_positions, fitted, cores, dt),
            "positions_tested_per_s": n_positions / dt, "seconds": dt}


# ------------------------------------------------------------------------------------------
# CPU arm, reference flavour: the unmodified reference CLI (oracle/_ref) on real files
# ------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(os.path.join(REPO, "oracle", "_ref", "xpore", "scripts", "xpore.py"))


def scratch_dir(prefix):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix=prefix, dir=base)


def runs_spec(spec):
    """The spec grouped by run, as the files are written (pooling is a YAML switch)."""
    from xpore_b200.synth import Spec
    m = {k: v for k, v in spec.method.items() if k != "pooling"}
    return Spec(spec.conditions, spec.ko_like, spec.n_genes, spec.pos_per_gene, spec.reads_lo, spec.reads_hi, spec.f_mod,
                spec.seed, spec.loguniform_total, m, spec.criteria)


def write_files(spec, km, n_genes, root):
    """Real data.json / data.index / config.yml of the first n_genes genes of the workload.  Returns
    (config path, run-grouped batch, bytes of data.json text)."""
    from xpore_b200.synth import make_batch, write_dataset_from_batch
    rs = runs_spec(spec)
    batch = make_batch(rs, km, n_positions=n_genes * spec.pos_per_gene, keys=True)
    full = type(spec)(spec.conditions, spec.ko_like, spec.n_genes, spec.pos_per_gene, spec.reads_lo, spec.reads_hi,
                      spec.f_mod, spec.seed, spec.loguniform_total, spec.method, spec.criteria)
    cfg = write_dataset_from_batch(root, full, batch, km, n_workers=len(os.sched_getaffinity(0)))
    nbytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs if f == "data.json")
    return cfg, batch, nbytes


def count_fitted_numpy(batch, spec):
    """Positions the prefilter gate lets through (diffmod.py:52-58), per gene id, computed with numpy/scipy on the
    run-grouped batch - only to COUNT the fitted positions of a reference run, whose table holds the filtered rows only."""
    import scipy.stats
    pre = spec.method.get("prefiltering", False)
    P = batch.n_positions
    sel = np.ones(P, dtype=bool)
    gc = np.asarray(batch.layout.run_cond)
    if pre and pre.get("method") == "t-test" and len(set(gc.tolist())) == 2:
        y = batch.y[:batch.n_reads]
        cond_of_read = np.repeat(np.tile(gc, P), batch.grp_len.reshape(-1))
        pos_of_read = np.repeat(np.arange(P), np.diff(batch.read_off))
        n = np.zeros((P, 2))
        s1 = np.zeros((P, 2))
        np.add.at(n, (pos_of_read, cond_of_read), 1.0)
        np.add.at(s1, (pos_of_read, cond_of_read), y)
        mean = s1 / np.maximum(n, 1)
        ss = np.zeros((P, 2))
        np.add.at(ss, (pos_of_read, cond_of_read), (y - mean[pos_of_read, cond_of_read]) ** 2)
        df = n.sum(axis=1) - 2
        svar = ss.sum(axis=1) / df
        t = (mean[:, 0] - mean[:, 1]) / np.sqrt(svar * (1 / n[:, 0] + 1 / n[:, 1]))
        p = scipy.stats.t.sf(np.abs(t), df) * 2
        valid = (n > 0).all(axis=1)
        sel = ~valid | np.isnan(p) | (p < pre["threshold"])
    by_id = {}
    for i, s in zip(batch.ids, sel):
        by_id[i] = by_id.get(i, 0) + int(s)
    return by_id


def run_reference_cli(cfg_path, ids, cores, timeout=1500):
    """Wall clock of the unmodified reference's `xpore diffmod --n_processes cores --ids ...` (interpreter start,
    imports and file reads included - what a user of the reference waits for)."""
    out_dir = os.path.join(os.path.dirname(cfg_path), "out")
    shutil.rmtree(out_dir, ignore_errors=True)
    cmd = [sys.executable, os.path.join(REPO, "oracle", "run_ref.py"), "diffmod", "--config", cfg_path, "--n_processes",
           str(cores), "--ids"] + list(ids)
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference CLI failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    log = open(os.path.join(out_dir, "diffmod.log")).read().splitlines()
    assert log[-1].startswith("--- SUCCESSFULLY"), log[-3:]
    rows = sum(1 for _ in open(os.path.join(out_dir, "diffmod.table"))) - 1
    return dt, rows


def reference_genes_for(spec, cores, seconds=8.0):
    """Genes (100 positions each) that keep one reference run near `seconds` on `cores` processes (survey: ~25
    fitted positions/s/core, ~600 tested/s/core through the t-test, ~0.7 pool efficiency)."""
    per_pos = float(np.mean([spec.reads_lo, spec.reads_hi])) * sum(spec.conditions.values()) if not spec.loguniform_total \
        else 900.0
    fit_s = 0.04 * max(per_pos / 360.0, 0.3) * (2.0 if len(spec.conditions) > 2 else 1.0)
    frac = 0.19 if (spec.method.get("prefiltering") and len(spec.conditions) == 2) else 1.0
    per_gene = spec.pos_per_gene * (0.0017 + frac * fit_s)
    return int(max(cores, min(spec.n_genes, round(seconds * cores * 0.7 / per_gene / cores) * cores)))


def reference_arm(args, spec, km, cores, workload):
    steps, warm = max(1, args.steps), max(0, args.warmup)
    if not reference_available():
        from xpore_b200.synth import make_batch
        n_cpu = args.cpu_positions or min(200000, 1600 * cores)  # ~10 s per step on the box's cores
        batch = make_batch(spec, km, n_positions=n_cpu)
        for _ in range(warm):
            cpu_baseline_port(batch, spec, min(batch.n_positions, 4 * cores), cores)
        vals = [cpu_baseline_port(batch, spec, batch.n_positions, cores) for _ in range(steps)]
        v = float(np.mean([x["value"] for x in vals]))
        ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
        cb = dict(vals[-1], value=v)
        sample = batch.n_positions
    else:
        n_genes = (args.cpu_positions // spec.pos_per_gene) if args.cpu_positions else reference_genes_for(spec, cores)
        root = scratch_dir("xpb200_ref_")
        try:
            cfg, batch, nbytes = write_files(spec, km, n_genes, root)
            ids = sorted(set(batch.ids))
            fitted = sum(count_fitted_numpy(batch, spec).values())
            for _ in range(min(warm, 1)):  # page cache, .pyc files
                run_reference_cli(cfg, ids[:cores], cores)
            walls = [run_reference_cli(cfg, ids, cores)[0] for _ in range(steps)]
        finally:
            shutil.rmtree(root, ignore_errors=True)
        wall = float(np.mean(walls))
        v, ms, sample = fitted / wall, wall * 1e3, batch.n_positions
        cb = {"value": v, "unit": "positions fitted/s", "cores": cores, "kind": "reference",
              "sample": "the unmodified reference CLI (oracle/_ref = a verbatim copy of /root/reference/xpore; import shims "
                        "ujson->json and an empty h5py) `xpore diffmod --n_processes %d` on real data.json/data.index files "
                        "of the first %d genes of the workload: %d positions, %d fitted, %.1f MB of text, %.1f s wall per run "
                        "(interpreter start and imports included)" % (cores, n_genes, sample, fitted, nbytes / 1e6, wall),
              "positions_tested_per_s": sample / wall, "seconds": wall}
    print(json.dumps({
        "impl": "reference", "metric": "diffmod positions fitted/sec", "value": v, "unit": "positions fitted/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sample_positions": sample, "host_cores": cores},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "positions fitted/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------
# The drop-in CLI, wall clock, on real files
# ------------------------------------------------------------------------------------------
def cli_e2e(spec, km, cores, local_rank, n_genes=None, with_reference=True):
    from xpore_b200.synth import make_batch  # noqa: F401
    per_pos = float(np.mean([spec.reads_lo, spec.reads_hi])) * sum(spec.conditions.values()) if not spec.loguniform_total \
        else 900.0
    # ~1.4e8 reads of text (c2: 4000 genes = 400k positions, 0.95 GB of data.json): large enough that interpreter start,
    # imports and the CUDA context (about 1 s together) do not decide the rate
    n_genes = n_genes or int(max(10, min(spec.n_genes, round(144e6 / per_pos / spec.pos_per_gene))))
    root = scratch_dir("xpb200_cli_")
    out = {}
    try:
        t0 = time.perf_counter()
        cfg, batch, nbytes = write_files(spec, km, n_genes, root)
        out["dataset"] = {"genes": n_genes, "positions": batch.n_positions, "reads": batch.n_reads,
                          "text_mb": nbytes / 1e6, "write_s": time.perf_counter() - t0, "dir": os.path.dirname(root)}
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
        runs = []
        tpath = os.path.join(root, "timing.json")
        for threads in sorted({1, max(1, cores // 2), cores}):
            shutil.rmtree(os.path.join(root, "out"), ignore_errors=True)
            cmd = [sys.executable, "-m", "xpore_b200.cli", "diffmod", "--config", cfg, "--n_processes", str(threads),
                   "--timing", tpath]
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=REPO, env=env, timeout=1500)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError("xpore_b200 CLI failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
            tm = json.load(open(tpath))
            rows = sum(1 for _ in open(os.path.join(root, "out", "diffmod.table"))) - 1
            runs.append({"ingest_threads": threads, "wall_s": wall, "in_process_s": tm["total_s"],
                         "positions_tested_per_s": tm["positions"] / wall, "positions_fitted_per_s": tm["fitted"] / wall,
                         "ingest_mb_per_s": tm["text_bytes"] / 1e6 / max(tm["ingest_s"], 1e-9),
                         "stages_s": {"ingest": tm["ingest_s"], "gpu_call": tm["gpu_s"], "rows": tm["rows_s"],
                                      "write": tm["write_s"]},
                         "gpu_call_ms_per_chunk": tm["gpu_s"] * 1e3 / max(tm["chunks"], 1), "chunks": tm["chunks"],
                         "gpu_call_ms_per_chunk_after_first": (tm["gpu_s"] - tm.get("gpu_first_chunk_s", 0.0)) * 1e3
                         / max(tm["chunks"] - 1, 1) if tm["chunks"] > 1 else None,
                         "cuda_init_s": tm.get("cuda_init_s"), "cuda_init_wait_s": tm.get("cuda_init_wait_s"),
                         "ingest_parse_s": tm.get("ingest_parse_s"),
                         "positions": tm["positions"], "fitted": tm["fitted"], "rows": rows})
        out["runs"] = runs
        best = min(runs, key=lambda x: x["wall_s"])
        out.update({"wall_s": best["wall_s"], "positions_fitted_per_s": best["positions_fitted_per_s"],
                    "positions_tested_per_s": best["positions_tested_per_s"], "ingest_threads": best["ingest_threads"],
                    "api": "python -m xpore_b200.cli diffmod --config <yaml> --n_processes <threads> (subprocess, wall clock "
                           "around it: interpreter start, imports, CUDA context, data.json -> diffmod.table)"})
        table_ours = open(os.path.join(root, "out", "diffmod.table")).read().splitlines()
        if with_reference and reference_available():
            g_ref = min(n_genes, reference_genes_for(spec, cores, seconds=10.0))
            ids = sorted(set(batch.ids))[:g_ref]
            wall, rows = run_reference_cli(cfg, ids, cores)
            by_id = count_fitted_numpy(batch, spec)
            fitted = sum(by_id[i] for i in ids)
            ref_keys = {tuple(l.split(",")[:3]) for l in open(os.path.join(root, "out", "diffmod.table")).read().splitlines()[1:]}
            our_keys = {tuple(l.split(",")[:3]) for l in table_ours[1:] if l.split(",")[0] in set(ids)}
            out["reference"] = {"kind": "reference", "cores": cores, "genes": g_ref, "positions": g_ref * spec.pos_per_gene,
                                "fitted": fitted, "wall_s": wall, "rows": rows,
                                "positions_fitted_per_s": fitted / wall,
                                "positions_tested_per_s": g_ref * spec.pos_per_gene / wall,
                                "same_rows_as_xpore_b200": ref_keys == our_keys,
                                "how": "the unmodified reference CLI (oracle/_ref) with --n_processes %d --ids <first %d ids> on "
                                       "the same files" % (cores, g_ref)}
            out["speedup_positions_per_s"] = best["positions_tested_per_s"] / out["reference"]["positions_tested_per_s"]
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


class Shard:
    """One rank's batch, resident in HBM and mirrored in pinned host memory, with everything a step needs."""

    def __init__(self, batch, method, dev, rank, world, local_rank, y_format_auto=True, index_offset=0):
        import ctypes as C

        import torch

        from xpore_b200 import _lib
        from xpore_b200 import fit as xf
        from xpore_b200.dist import RowGatherer
        self.torch, self.xf, self.L, self.C, self._lib = torch, xf, _lib.lib(), C, _lib
        self.rank, self.world, self.local_rank, self.method, self.dev = rank, world, local_rank, method, dev
        self.yfmt = batch.encode_reads() if y_format_auto else 0
        self.db = xf.DeviceBatch(batch, device=dev, pin=True)
        self.out = xf.DeviceResult(self.db.P, self.db.G, self.db.C, dev)
        self.n_reads = torch.from_numpy(np.diff(batch.read_off)).to(dev)
        self.fit_w, self.stats_w = xf.fit_width(self.db.G), xf.stats_width(self.db.G, self.db.C)
        self.rec_w = 4 + self.fit_w + self.stats_w
        self.index_offset = int(index_offset)
        hb = self.db.host
        self.cb = _lib.XpBatch(self.db.P, self.db.G, self.db.C, self.db.max_reads, self.db.R, hb["y"].data_ptr(),
                               hb["read_off"].data_ptr(), hb["grp_len"].data_ptr(), hb["kmer_mean"].data_ptr(),
                               hb["kmer_std"].data_ptr(), hb["grp_cond"].data_ptr(), self.db.y_format, 0, self.db.y_scale)
        self.cm = method.c_struct()
        # row capacity of the gather: from one untimed pass (every rank), the largest count over the ranks + 25 %
        xf.fit_device(self.db, method, self.out)
        torch.cuda.synchronize()
        self.n_fitted = int(self.out.n_fitted.item())
        keep = int(((self.out.status & (xf.SELECTED | xf.KEEP_ROW)) == (xf.SELECTED | xf.KEEP_ROW)).sum().item())
        cap = torch.tensor([keep], dtype=torch.int64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        self.cap = int(cap.item()) + int(cap.item()) // 4 + 64
        self.gather = RowGatherer(self.cap, self.rec_w, rank, world, dev)
        self.d2h = 0
        self.hrows = torch.empty((max(1, self.cap), self.rec_w), dtype=torch.float64).pin_memory() if world == 1 else None
        self.nrows, self.nf = C.c_int64(0), C.c_uint32(0)

    def step_device(self):
        """HBM-resident inputs: fit; N > 1: + the table's rows compacted on the device + NCCL gather to rank 0."""
        C, xf = self.C, self.xf
        xf.fit_device(self.db, self.method, self.out)
        if self.world > 1:
            r = self.out.c_struct()
            stream = self.torch.cuda.current_stream().cuda_stream
            self._lib.check(self.L.xpb200_pack_rows(C.byref(r), self.db.P, self.db.G, self.db.C, self.out.workspace.data_ptr(),
                                                    self.index_offset, self.gather.rows.data_ptr(), self.cap,
                                                    self.gather.counts.data_ptr(), C.c_void_p(stream)))
            self.gather.gather()

    def step_e2e(self):
        """Host buffers through the C ABI.  N = 1: xpb200_fit_host_rows (rows to pinned host memory).  N > 1:
        xpb200_fit_host_rows_async (rows stay on the device, no host synchronisation) -> NCCL gather -> rank 0 copies
        exactly the rows to its pinned host memory."""
        C = self.C
        if self.world == 1:
            self._lib.check(self.L.xpb200_fit_host_rows(self.local_rank, C.byref(self.cb), C.byref(self.cm), self.hrows.data_ptr(),
                                                        self.cap, C.byref(self.nrows), C.byref(self.nf)))
            self.d2h = int(self.nrows.value) * self.rec_w * 8 + 4 * 16
            return
        stream = self.torch.cuda.current_stream().cuda_stream
        self._lib.check(self.L.xpb200_fit_host_rows_async(self.local_rank, C.byref(self.cb), C.byref(self.cm),
                                                          self.gather.rows.data_ptr(), self.cap, self.index_offset,
                                                          self.gather.counts.data_ptr(), C.c_void_p(stream)))
        self.gather.gather()
        if self.rank == 0:
            _, _, self.d2h = self.gather.fetch()


def timed(shard, fn, steps, warmup, sync_all, wall=False):
    """ms per step of fn over `steps` steps after `warmup`: CUDA events on the current stream (or wall clock around a
    host-synchronising call), barrier + synchronize on both sides."""
    torch = shard.torch
    for _ in range(warmup):
        fn()
    sync_all()
    if wall:
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        sync_all()
        return (time.perf_counter() - t0) * 1e3 / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync_all()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--positions", type=int, default=None, help="positions per GPU (default: the config's)")
    ap.add_argument("--cpu-positions", type=int, default=None, help="positions in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the cli_e2e leg (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement (N > 1)")
    ap.add_argument("--cli-genes", type=int, default=None, help="genes (x100 positions) of the cli_e2e dataset")
    ap.add_argument("--y-format", default="auto", choices=["auto", "f64"],
                    help="auto: ship reads as scaled u16/i32 when that is bit-exact (2-decimal dataprep output)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from xpore_b200.kmer_model import KmerModel
    from xpore_b200.synth import config_spec

    spec = config_spec(args.config)
    P_cfg = spec.n_genes * spec.pos_per_gene
    P = args.positions or P_cfg
    km = KmerModel.load()
    workload = WORKLOADS[args.config]
    cores = len(os.sched_getaffinity(0))

    # ---------------------------------------------------------------- reference (CPU) arm
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, spec, km, cores, workload)
        return

    # ---------------------------------------------------------------- B200 arm
    # CPU legs first (rank 0 at N = 1 only), before CUDA is initialised in this process (fork-based pools)
    cpu = None
    cli = None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            if reference_available():
                n_genes = (args.cpu_positions // spec.pos_per_gene) if args.cpu_positions else reference_genes_for(spec, cores, 12.0)
                root = scratch_dir("xpb200_ref_")
                try:
                    cfg, rb, nbytes = write_files(spec, km, n_genes, root)
                    ids = sorted(set(rb.ids))
                    fitted = sum(count_fitted_numpy(rb, spec).values())
                    wall, _ = run_reference_cli(cfg, ids, cores)
                finally:
                    shutil.rmtree(root, ignore_errors=True)
                cpu = {"value": fitted / wall, "unit": "positions fitted/s", "cores": cores, "kind": "reference",
                       "sample": "the unmodified reference CLI (oracle/_ref) `xpore diffmod --n_processes %d` on real files of "
                                 "the first %d genes of the workload: %d positions, %d fitted, %.1f s wall"
                                 % (cores, n_genes, rb.n_positions, fitted, wall),
                       "positions_tested_per_s": rb.n_positions / wall, "seconds": wall}
            else:
                from xpore_b200.synth import make_batch
                n_cpu = args.cpu_positions or min(200000, 2400 * cores)  # ~15 s of CPU work
                cb_batch = make_batch(spec, km, n_positions=n_cpu)
                cpu = cpu_baseline_port(cb_batch, spec, cb_batch.n_positions, cores)
                del cb_batch
        if not args.no_cli:
            try:
                cli = cli_e2e(spec, km, cores, local_rank, n_genes=args.cli_genes)
            except Exception as e:  # the hot-path numbers below do not depend on this leg
                cli = {"error": repr(e)[:500]}

    import ctypes as C

    import torch
    import torch.distributed as dist

    from xpore_b200 import _lib
    from xpore_b200 import fit as xf
    from xpore_b200.synth import make_batch_device

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bound_cpus = 0
    if world > 1:  # before any pinned allocation: host buffers next to this rank's GPU
        from xpore_b200.dist import bind_to_gpu_cpus
        bound_cpus = bind_to_gpu_cpus(local_rank)
        dist.init_process_group("nccl", device_id=dev)
    _lib.build()
    L = _lib.lib()

    peak = C.c_double(0.0)
    _lib.check(L.xpb200_measure_fp64_peak(local_rank, 5, C.byref(peak)))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    method = xf.Method.from_dict(spec.method)
    batch = make_batch_device(spec, dev, km, n_positions=P, seed_offset=rank)
    sh = Shard(batch, method, dev, rank, world, local_rank, args.y_format == "auto", index_offset=rank * P)
    db, out = sh.db, sh.out
    del batch

    # ---- value: HBM-resident inputs ------------------------------------------------------
    for _ in range(args.warmup):
        sh.step_device()
    sync_all()
    launches0 = L.xpb200_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(sh, sh.step_device, args.steps, 0, sync_all) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    launches = L.xpb200_launch_count() - launches0
    # per-stage durations of further (untimed) steps, CUDA events on the launch stream
    stage_ms = np.zeros((args.steps, 4))
    _lib.check(L.xpb200_set_stage_timing(1))
    for i in range(args.steps):
        xf.fit_device(db, method, out)
        ms4 = (C.c_float * 4)()
        _lib.check(L.xpb200_get_stage_ms(ms4))
        stage_ms[i] = list(ms4)
    _lib.check(L.xpb200_set_stage_timing(0))
    torch.cuda.synchronize()

    n_fitted = int(out.n_fitted.item())
    sel = (out.status & 1) != 0
    iters = out.n_iter.to(torch.float64)
    read_iters = float((sh.n_reads.to(torch.float64) * iters * sel).sum().item())
    pos_iters = float((iters * sel).sum().item())
    sel_reads = float((sh.n_reads * sel).sum().item())
    it_sel = out.n_iter[sel].float()

    def reduce_ms_and_counts(ms, fitted, tested):
        t = torch.tensor([ms, float(fitted), float(tested)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tsum[1]), float(tsum[2])
        return ms, float(fitted), float(tested)

    total_ms, fitted_all, tested_all = reduce_ms_and_counts(total_ms, n_fitted, db.P)
    ms_per_step = total_ms / args.steps
    value = fitted_all / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI ----------------------------------------------
    e2e_ms = timed(sh, sh.step_e2e, args.steps, max(1, min(args.warmup, 2)), sync_all, wall=True)
    e2e_ms, _, _ = reduce_ms_and_counts(e2e_ms, 0, 0)
    e2e_value = fitted_all / (e2e_ms * 1e-3)
    fit_w, stats_w = sh.fit_w, sh.stats_w
    h2d_bytes, d2h_bytes, yfmt = db.h2d_bytes, sh.d2h, sh.yfmt
    keep = ((out.status.to(torch.int64) & (xf.SELECTED | xf.KEEP_ROW)) == (xf.SELECTED | xf.KEEP_ROW)).nonzero().flatten()
    if world == 1:
        # the rows that came back are the device-resident run's rows, record for record
        assert int(sh.nf.value) == n_fitted
        got = sh.hrows[:int(sh.nrows.value)].to(dev)
        assert got.shape[0] == keep.numel()
        assert (got[:, 0].to(torch.int64) == keep).all()
        assert (got[:, 2].to(torch.int32) == out.n_iter[keep]).all()
        assert torch.equal(got[:, 4:4 + fit_w], out.fit[keep])
    elif rank == 0:
        host_rows, counts, _ = sh.gather.fetch()
        assert counts[0][0] == keep.numel() and counts[0][1] == n_fitted, (counts[0], keep.numel(), n_fitted)
        mine = host_rows[:counts[0][0]].to(dev)
        assert (mine[:, 0].to(torch.int64) == keep).all() and torch.equal(mine[:, 4:4 + fit_w], out.fit[keep])
        assert sum(c[1] for c in counts) == int(fitted_all)

    # ---- strong scaling (N > 1): the config's batch cut by the coverage-balanced partition ----------
    strong = None
    if world > 1 and not args.no_strong:
        from xpore_b200.dist import balanced_partition
        del sh
        torch.cuda.empty_cache()
        full = make_batch_device(spec, dev, km, n_positions=P, seed_offset=0)  # the same batch on every rank
        parts = balanced_partition(np.diff(full.read_off).astype(np.float64), world)
        mine = full.subset(parts[rank])
        reads_mine = int(mine.n_reads)
        del full
        ss = Shard(mine, method, dev, rank, world, local_rank, args.y_format == "auto", index_offset=int(parts[rank][0]) if len(parts[rank]) else 0)
        s_dev = timed(ss, ss.step_device, args.steps, args.warmup, sync_all)
        s_e2e = timed(ss, ss.step_e2e, args.steps, 2, sync_all, wall=True)
        s_dev, s_fit, s_tested = reduce_ms_and_counts(s_dev, ss.n_fitted, ss.db.P)
        s_e2e, _, _ = reduce_ms_and_counts(s_e2e, 0, 0)
        rd = torch.tensor([float(reads_mine)], dtype=torch.float64, device=dev)
        rmax, rsum = rd.clone(), rd.clone()
        dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(rsum, op=dist.ReduceOp.SUM)
        strong = {"scaling": "strong", "positions_total": int(s_tested), "positions_fitted_total": int(s_fit),
                  "partition": "xpore_b200.dist.balanced_partition: contiguous ranges of ~equal read count",
                  "max_over_mean_reads_per_rank": float(rmax[0] / (rsum[0] / world)),
                  "ms_per_step": s_dev, "value": s_fit / (s_dev * 1e-3),
                  "e2e_ms_per_step": s_e2e, "e2e_value": s_fit / (s_e2e * 1e-3)}
        sh = ss

    if rank == 0:
        fit_ms = float(np.median(stage_ms[:, 2]))
        G = db.G
        flops = read_iters * FLOP_PER_READ_ITER + pos_iters * (6 * G + 4) * FLOP_PER_SPECIAL
        achieved = flops / (fit_ms * 1e-3) / 1e12
        ybytes = {0: 8, 1: 4, 2: 2}[yfmt]
        alg_bytes = sel_reads * ybytes + n_fitted * (8 + 4 * G + 16 + 4) + n_fitted * (fit_w * 8 + 8)
        # DRAM traffic of the fit kernel per launch from the committed ncu capture of this very workload
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(REPO, "profiles", "fit_traffic.json"))).get(args.config)
            yname = {0: "f64", 1: "i32 scaled by 100", 2: "u16 scaled by 100"}[yfmt]
            if tr and tr["positions_per_gpu"] == db.P and tr["y_format"] == yname:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                traffic_src = "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch (%s)" % tr["source"]
        except Exception:
            pass
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        info = _lib.XpLaunchInfo()
        _lib.check(L.xpb200_fit_launch_info(local_rank, db.G, db.max_reads, db.y_format, C.byref(info)))
        theo = FP64_THEORETICAL(info.sm_count, (clocks or {}).get("sm_max_mhz") or 1965.0)
        pf_ms = float(np.median(stage_ms[:, 0]))
        pf_bytes = db.R * ybytes + db.P * (8 + 4 * G + 8 + 12)
        line = {
            "metric": "diffmod positions fitted/sec", "value": value, "unit": "positions fitted/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "positions_per_gpu": db.P, "reads_per_gpu": db.R,
                       "positions_fitted_per_gpu": n_fitted, "groups": db.G, "conditions": db.C,
                       "max_iters": method.max_iters, "prefilter": method.prefilter_method,
                       "pooling": method.pooling, "parallelism": "positions sharded, dp%d" % world,
                       "y_format": {0: "f64", 1: "i32 scaled by 100", 2: "u16 scaled by 100"}[yfmt],
                       "l2": "inputs (%.2f GB per GPU) larger than the 126 MB L2; no explicit flush"
                             % (db.h2d_bytes / 1e9),
                       "sweeps_mean": float(it_sel.mean().item()) if n_fitted else 0.0,
                       "sweeps_max": int(it_sel.max().item()) if n_fitted else 0,
                       "positions_tested_per_s": tested_all / (ms_per_step * 1e-3),
                       "read_sweeps_per_s": read_iters * world / (ms_per_step * 1e-3),
                       "scaling_detail": strong},
            "e2e": {"value": e2e_value, "unit": "positions fitted/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "xpb200_fit_host_rows (C ABI, pinned host buffers in, records of the rows of diffmod.table out)"
                           if world == 1 else
                           "xpb200_fit_host_rows_async per rank (pinned host buffers in, the table's rows compacted on the "
                           "device, no host synchronisation) -> NCCL gather of fixed-capacity row buffers -> rank 0 copies "
                           "exactly the rows to pinned host memory"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "xp_fit_tm_kernel" if info.tmem_columns else "xp_fit_kernel", "bound": "fp64", "achieved": achieved, "peak": peak.value,
                         "unit": "TFLOP/s", "frac": achieved / peak.value,
                         "peak_source": "measured here: xpb200_measure_fp64_peak (8 independent DFMA chains per "
                                        "thread, 8 CTAs x 256 threads per SM, best of 5); committed record with clocks: "
                                        "profiles/r2_fp64_peak.json",
                         "peak_theoretical": theo, "frac_of_theoretical": achieved / theo,
                         "flop_model": "100 flop per read-sweep + (6G+4)*60 per position-sweep (SURVEY.md 8d): the survey's "
                                       "NOMINAL cost of the algorithm, not the instructions executed (the kernel spends 28 "
                                       "FP64 instructions per read-sweep) - frac is a throughput against the nominal model, "
                                       "not a pipe utilisation (ncu: profiles/)",
                         "ms": fit_ms, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes,
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (fit_ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (fit_ms * 1e-3) / 1e9 / hbm_peak}},
            "stages_ms": {"prefilter": pf_ms, "worklist": float(np.median(stage_ms[:, 1])),
                          "fit": fit_ms, "stats": float(np.median(stage_ms[:, 3]))},
            "prefilter_hbm": {"bytes": pf_bytes, "achieved_gbs": pf_bytes / (pf_ms * 1e-3) / 1e9 if pf_ms > 0 else None,
                              "peak_gbs": hbm_peak, "frac": pf_bytes / (pf_ms * 1e-3) / 1e9 / hbm_peak if pf_ms > 0 else None},
            "launch": {"sm_count": info.sm_count, "ctas_per_sm": info.ctas_per_sm, "warps_per_cta": info.warps_per_cta,
                       "smem_per_cta": info.smem_per_cta, "slab_reads": info.slab_reads,
                       "regs_per_thread": info.regs_per_thread, "positions_per_warp": info.positions_per_warp,
                       "tmem_columns_per_position": info.tmem_columns,
                       "kernel": "xp_fit_tm_kernel (reads in tensor memory)" if info.tmem_columns else "xp_fit_kernel (reads in shared memory)"},
            "cpu_baseline": cpu,
            "cli_e2e": cli,
            "host_cores": cores, "cpus_bound_to_gpu": bound_cpus,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
