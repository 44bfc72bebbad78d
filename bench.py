#!/usr/bin/env python
"""Benchmark of the diffmod hot path: positions fitted per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--positions P]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle port on all host cores

One step = one pass of the whole path (prefilter t-test -> worklist -> VB-EM fit -> statistics)
over one synthetic batch of the BASELINE.json shape (default c2: 2 conditions x 3 replicates,
1M positions, 20-100 reads per run, t-test prefilter at 0.1).

`value`  positions fitted/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    the same metric through the C ABI with HOST buffers (`xpb200_fit_host`): pinned-host ->
         device copies of the batch and device -> host copies of the result records inside the
         timed region.
`roofline`  the fit kernel against the measured FP64 DFMA peak (this path is FP64-pipe bound,
         SURVEY.md §8d), plus its HBM figures.
`cpu_baseline`  the numpy/scipy oracle port (same per-position call structure as the reference)
         on all host cores, on a bounded sample of the same batch.
Multi-GPU: each rank fits its own shard (weak scaling), results of fitted positions are gathered
to rank 0 over NCCL inside the timed step.
"""
import argparse
import json
import os
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


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port, one task = a chunk of positions (the reference's task = one gene)
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


def cpu_baseline(batch, spec, n_positions, cores=None, chunk=25):
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
    ap.add_argument("--y-format", default="auto", choices=["auto", "f64"],
                    help="auto: ship reads as scaled u16/i32 when that is bit-exact (2-decimal dataprep output)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from xpore_b200.kmer_model import KmerModel
    from xpore_b200.synth import config_spec, make_batch

    spec = config_spec(args.config)
    P_cfg = spec.n_genes * spec.pos_per_gene
    P = args.positions or P_cfg
    km = KmerModel.load()
    workload = {"c1": "c1: 1 gene x 100 positions, 2x2, 50-200 reads/run, t-test prefilter 0.1",
                "c2": "c2: HEK293T WT vs METTL3-KO shape, 2 conditions x 3 replicates, 1M positions, "
                      "20-100 reads/run, synthetic reads, t-test prefiltering 0.1",
                "c3": "c3: c2 shape, pooling=True, no prefilter, max_iters=500",
                "c4": "c4: 100k positions, 2x2, 20-5000 reads/position log-uniform",
                "c5": "c5: 6 conditions x 3 replicates, 2M positions, 20-60 reads/run"}[args.config]
    cores = len(os.sched_getaffinity(0))

    # ---------------------------------------------------------------- reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = args.cpu_positions or min(200000, 1600 * cores)  # ~10 s per step on the box's cores
        batch = make_batch(spec, km, n_positions=n_cpu)
        n_cpu = batch.n_positions
        vals = []
        for _ in range(args.warmup):
            cpu_baseline(batch, spec, min(n_cpu, 4 * cores), cores)
        for _ in range(max(1, args.steps)):
            vals.append(cpu_baseline(batch, spec, n_cpu, cores))
        v = float(np.mean([x["value"] for x in vals]))
        ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
        cb = dict(vals[-1], value=v)
        print(json.dumps({
            "impl": "reference", "metric": "diffmod positions fitted/sec", "value": v, "unit": "positions fitted/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample_positions": n_cpu, "host_cores": cores},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "positions fitted/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ---------------------------------------------------------------- B200 arm
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (fork-based pool)
        n_cpu = args.cpu_positions or min(200000, 2400 * cores)  # ~15 s of CPU work
        cb_batch = make_batch(spec, km, n_positions=n_cpu)
        cpu = cpu_baseline(cb_batch, spec, cb_batch.n_positions, cores)
        del cb_batch

    import torch
    import torch.distributed as dist

    from xpore_b200 import _lib
    from xpore_b200 import fit as xf
    from xpore_b200.dist import gather_records
    from xpore_b200.synth import make_batch_device

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bound_cpus = 0
    if world > 1:  # before any pinned allocation: host buffers next to this rank's GPU
        from xpore_b200.dist import bind_to_gpu_cpus
        bound_cpus = bind_to_gpu_cpus(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.build()
    L = _lib.lib()
    import ctypes as C

    peak = C.c_double(0.0)
    _lib.check(L.xpb200_measure_fp64_peak(local_rank, 5, C.byref(peak)))

    batch = make_batch_device(spec, dev, km, n_positions=P, seed_offset=rank)
    method = xf.Method.from_dict(spec.method)
    yfmt = batch.encode_reads() if args.y_format == "auto" else 0
    db = xf.DeviceBatch(batch, device=dev, pin=True)
    out = xf.DeviceResult(db.P, db.G, db.C, dev)
    torch.cuda.synchronize()
    n_reads = torch.from_numpy(np.diff(batch.read_off)).to(dev)
    fit_w, stats_w = xf.fit_width(db.G), xf.stats_width(db.G, db.C)

    recbuf = [None]

    def step_device():
        xf.fit_device(db, method, out)
        if world > 1:
            rec, recbuf[0] = xf.pack_records_device(out, index_offset=rank * db.P, buf=recbuf[0])
            return gather_records(rec, rank, world)
        return None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident inputs ------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    _lib.check(L.xpb200_set_stage_timing(1))
    stage_ms = np.zeros((args.steps, 4))
    launches0 = L.xpb200_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    ev[0].record()
    for i in range(args.steps):
        step_device()
        ev[i + 1].record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = L.xpb200_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    _lib.check(L.xpb200_set_stage_timing(0))
    # per-stage durations of one more (untimed) step, CUDA events on the launch stream
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
    read_iters = float((n_reads.to(torch.float64) * iters * sel).sum().item())
    pos_iters = float((iters * sel).sum().item())
    sel_reads = float((n_reads * sel).sum().item())
    it_sel = out.n_iter[sel].float()

    t = torch.tensor([total_ms, float(n_fitted), float(db.P)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, fitted_all, tested_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        fitted_all, tested_all = float(n_fitted), float(db.P)
    ms_per_step = total_ms / args.steps
    value = fitted_all / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI ----------------------------------------------
    hb = db.host
    cb = _lib.XpBatch(db.P, db.G, db.C, db.max_reads, db.R, hb["y"].data_ptr(), hb["read_off"].data_ptr(),
                      hb["grp_len"].data_ptr(), hb["kmer_mean"].data_ptr(), hb["kmer_std"].data_ptr(),
                      hb["grp_cond"].data_ptr(), db.y_format, 0, db.y_scale)
    cm = method.c_struct()
    cr = _lib.XpResult(out.status.data_ptr(), out.n_iter.data_ptr(), out.pval.data_ptr(), out.fit.data_ptr(),
                       out.stats.data_ptr())  # N > 1: the dense results stay on the device of each rank
    nf = C.c_uint32(0)
    h2d = db.h2d_bytes
    d2h_box = [0]
    pinned_out = [None]

    # N = 1: xpb200_fit_host_rows - pinned host buffers in, the records of diffmod.table's rows (fitted positions
    # that pass the reference's row filter, io.py:363-367) out to pinned host memory: what `execute()` hands to
    # the table writer.  N > 1: xpb200_fit_host with device result arrays, the same rows compacted on the
    # device, gathered to rank 0 over NCCL and copied to its host there (the CLI gathers the same rows).
    rec_w = 4 + fit_w + stats_w
    rows_cap = max(1, n_fitted)
    hrows = torch.empty((rows_cap, rec_w), dtype=torch.float64).pin_memory() if world == 1 else None
    nrows = C.c_int64(0)

    def step_e2e():
        if world == 1:
            _lib.check(L.xpb200_fit_host_rows(local_rank, C.byref(cb), C.byref(cm), hrows.data_ptr(), rows_cap,
                                              C.byref(nrows), C.byref(nf)))
            d2h_box[0] = int(nrows.value) * rec_w * 8 + 4 * 16
        else:
            _lib.check(L.xpb200_fit_host(local_rank, C.byref(cb), C.byref(cm), C.byref(cr), C.byref(nf)))
            rec, recbuf[0] = xf.pack_records_device(out, index_offset=rank * db.P, buf=recbuf[0])
            rec = rec[(rec[:, 1].to(torch.int64) & xf.KEEP_ROW) != 0]
            allrec = gather_records(rec, rank, world)
            if rank == 0:
                if pinned_out[0] is None or pinned_out[0].shape[0] < allrec.shape[0]:
                    pinned_out[0] = torch.empty((int(allrec.shape[0] * 1.2) + 1, allrec.shape[1]),
                                                dtype=torch.float64).pin_memory()
                pinned_out[0][:allrec.shape[0]].copy_(allrec, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                d2h_box[0] = allrec.numel() * 8 + 4

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    e2e_value = fitted_all / (e2e_ms * 1e-3)
    if world == 1:
        assert int(nf.value) == n_fitted
        if not os.environ.get("XPB200_BENCH_NOVERIFY"):
            # the rows that came back are the device-resident run's rows, record for record
            keep = ((out.status.to(torch.int64) & (xf.SELECTED | xf.KEEP_ROW)) == (xf.SELECTED | xf.KEEP_ROW)).nonzero().flatten()
            got = hrows[:int(nrows.value)].to(dev)
            assert got.shape[0] == keep.numel()
            assert (got[:, 0].to(torch.int64) == keep).all()
            assert (got[:, 2].to(torch.int32) == out.n_iter[keep]).all()
            assert torch.equal(got[:, 4:4 + fit_w], out.fit[keep])

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
            tr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                                             "r1_fit_traffic.json"))).get(args.config)
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
        _lib.check(L.xpb200_fit_launch_info(local_rank, db.G, db.max_reads, C.byref(info)))
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
                       "read_sweeps_per_s": read_iters * world / (ms_per_step * 1e-3)},
            "e2e": {"value": e2e_value, "unit": "positions fitted/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_box[0],
                    "api": "xpb200_fit_host_rows (C ABI, pinned host buffers in, records of the rows of diffmod.table out)"
                           if world == 1 else
                           "xpb200_fit_host (pinned host buffers in, device records out) -> "
                           "xpb200_pack_records -> rows that pass the row filter -> NCCL gather -> rank-0 host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "xp_fit_kernel", "bound": "fp64", "achieved": achieved, "peak": peak.value,
                         "unit": "TFLOP/s", "frac": achieved / peak.value,
                         "peak_source": "measured here: xpb200_measure_fp64_peak (8 independent DFMA chains per "
                                        "thread, 8 CTAs x 256 threads per SM, best of 5)",
                         "flop_model": "100 flop per read-sweep + (6G+4)*60 per position-sweep (SURVEY.md 8d)",
                         "ms": fit_ms, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes,
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (fit_ms * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (fit_ms * 1e-3) / 1e9 / hbm_peak}},
            "stages_ms": {"prefilter": float(np.median(stage_ms[:, 0])), "worklist": float(np.median(stage_ms[:, 1])),
                          "fit": fit_ms, "stats": float(np.median(stage_ms[:, 3]))},
            "prefilter_hbm": {"bytes": db.R * ybytes + db.P * (8 + 4 * G + 8 + 12),
                              "achieved_gbs": (db.R * ybytes + db.P * (8 + 4 * G + 8 + 12))
                              / (float(np.median(stage_ms[:, 0])) * 1e-3) / 1e9 if stage_ms[:, 0].any() else None,
                              "peak_gbs": hbm_peak},
            "launch": {"sm_count": info.sm_count, "ctas_per_sm": info.ctas_per_sm, "warps_per_cta": info.warps_per_cta,
                       "smem_per_cta": info.smem_per_cta, "slab_reads": info.slab_reads,
                       "regs_per_thread": info.regs_per_thread},
            "cpu_baseline": cpu,
            "host_cores": cores, "cpus_bound_to_gpu": bound_cpus,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
