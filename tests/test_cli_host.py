"""Host side of the `diffmod` drop-in (config, ids, packing, table/log files, resume, ingest pool,
world_size-2 sharding over gloo).  The fit function is injected: the CPU suite plugs the host
emulator in as the checker; the product default is the CUDA path (tests/test_gpu_parity.py)."""
import argparse
import os

import numpy as np
import pytest

from xpore_b200 import diffmod as xd
from xpore_b200.batch import Layout
from xpore_b200.config import Configurator

from conftest import load_case
from helpers import assert_same_rows, compare_tables, emul_fit_batch, load_emulator, read_table


def emul_fit_fn(batch, method):
    return emul_fit_batch(load_emulator(), batch, method)


def run_cli(config_path, **kw):
    args = argparse.Namespace(config=config_path, n_processes=kw.get("n_processes", 1), save_models=False,
                              resume=kw.get("resume", False), ids=kw.get("ids", []),
                              batch_genes=kw.get("batch_genes", 8192), batch_bytes=kw.get("batch_bytes", 256 << 20))
    xd.diffmod(args, fit_fn=emul_fit_fn)
    out = Configurator(config_path).get_paths()["out_dir"]
    return os.path.join(out, "diffmod.table"), os.path.join(out, "diffmod.log")


def expected_rows(exp):
    return {tuple(r[:3]): r for r in exp["rows"]}


def test_ids_and_table_files(golden_case):
    exp = golden_case["expected"]
    table, log = run_cli(golden_case["config_path"])
    cfg = Configurator(golden_case["config_path"])
    method = cfg.get_method()
    lay = Layout.from_data_info(cfg.get_data_info(), method["pooling"])
    header, rows = read_table(table)
    assert header == exp["header"]
    pm = method["prefiltering"]["method"] if method["prefiltering"] else None
    compare_tables(rows, expected_rows(exp), header, lay, pm, rtol_param=1e-8, rtol_p=1e-8)
    lines = open(log).read().splitlines()
    assert lines[0] == "--- DIFFMOD ---" and lines[-1] == "--- SUCCESSFULLY FINISHED ---"
    assert lines[1:-1] == exp["ids"]  # helper.get_ids: sorted ids present in >= 2 conditions
    assert open(table, "rb").read().count(b"\r\n") == len(rows) + 1  # csv default dialect, like the reference


def test_resume_ids_small_batches_and_pool(tmp_path):
    case = load_case("g1_runs", tmp_path)
    exp = case["expected"]
    table, log = run_cli(case["config_path"], ids=[exp["ids"][0]])
    _, rows1 = read_table(table)
    assert {k[0] for k in rows1} <= {exp["ids"][0]}
    assert open(log).read().splitlines()[1:-1] == [exp["ids"][0]]
    # resume: the finished id is skipped, rows are appended, nothing is duplicated
    table, log = run_cli(case["config_path"], resume=True)
    header, rows = read_table(table)
    assert set(rows) == set(expected_rows(exp))
    assert len(open(table).read().splitlines()) == len(rows) + 1
    # tiny GPU batches and a 2-process ingest pool give the same table
    table2, _ = run_cli(case["config_path"], batch_genes=1, n_processes=2)
    _, rows2 = read_table(table2)
    assert rows2 == rows


def _rank_main(rank, world, config_path, port):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    run_cli(config_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["g1_runs", "g7_six_conditions"])
def test_two_ranks_gloo(tmp_path, name):
    """N > 1 path on CPU: ids split by data.index byte spans, the numeric row records gathered on rank 0 (one gene and
    18 group slots in the second case: a rank without work, NaN prefilter cells, absent-group masks)."""
    import torch.multiprocessing as mp
    case = load_case(name, tmp_path)
    table1, _ = run_cli(case["config_path"])
    _, rows1 = read_table(table1)
    os.rename(table1, table1 + ".single")
    mp.spawn(_rank_main, args=(2, case["config_path"], 29617), nprocs=2, join=True)
    out = Configurator(case["config_path"]).get_paths()["out_dir"]
    header, rows2 = read_table(os.path.join(out, "diffmod.table"))
    assert_same_rows(rows1, rows2)
    lines = open(os.path.join(out, "diffmod.log")).read().splitlines()
    assert sorted(lines[1:-1]) == case["expected"]["ids"] and lines[-1] == "--- SUCCESSFULLY FINISHED ---"


def _gather_main(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    from xpore_b200.dist import gather_records, pack_records
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, FW, SW = 10 + 3 * rank, 5, 4
    g = torch.Generator().manual_seed(rank)
    status = (torch.rand(P, generator=g) < 0.5).to(torch.int32) * 3
    n_iter = torch.arange(P, dtype=torch.int32) + 100 * rank
    pval = torch.rand(P, generator=g, dtype=torch.float64)
    fit = torch.rand((P, FW), generator=g, dtype=torch.float64)
    stats = torch.rand((P, SW), generator=g, dtype=torch.float64)
    rec = pack_records(status, n_iter, pval, fit, stats, index_offset=1000 * rank)
    allrec = gather_records(rec, rank, world)
    if rank == 0:
        ret["all"] = allrec.numpy()
    ret[rank] = rec.numpy()
    dist.destroy_process_group()


def test_record_gather_gloo():
    import torch.multiprocessing as mp

    from xpore_b200.dist import balanced_partition, unpack_records
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gather_main, args=(2, 29618, ret), nprocs=2, join=True)
        allrec = ret["all"]
        np.testing.assert_array_equal(allrec, np.concatenate([ret[0], ret[1]]))
        idx, status, n_iter, pval, fit, stats = unpack_records(allrec, 5, 4)
        assert (status & 1).all() and fit.shape[1] == 5 and stats.shape[1] == 4
        assert (idx[: len(ret[0])] < 1000).all() and (idx[len(ret[0]):] >= 1000).all()
    # coverage-balanced contiguous partition
    cost = np.array([5, 1, 1, 1, 1, 1, 10, 1, 1, 2], dtype=float)
    parts = balanced_partition(cost, 3)
    assert np.array_equal(np.concatenate(parts), np.arange(10))
    assert max(cost[p].sum() for p in parts) <= 12
    assert [len(p) for p in balanced_partition(np.ones(0), 2)] == [0, 0]
