import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
CASES = sorted(d for d in os.listdir(GOLDEN) if os.path.isfile(os.path.join(GOLDEN, d, "expected.json")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`@pytest.mark.gpu` tests are skipped (not failed) on a box without a CUDA device, unless they were asked
    for explicitly with `-m gpu` - there a missing device must fail loudly (the driver runs `-m gpu` on a B200)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_case(name, tmp_path=None):
    """Golden case -> dict(spec, expected, data_info, method, criteria[, config_path])."""
    root = os.path.join(GOLDEN, name)
    with open(os.path.join(root, "spec.json")) as f:
        spec = json.load(f)
    with open(os.path.join(root, "expected.json")) as f:
        exp = json.load(f)
    cfg = {"data": {c: {r: os.path.join(root, d) for r, d in runs.items()} for c, runs in spec["data"].items()}}
    if spec.get("method"):
        cfg["method"] = spec["method"]
    if spec.get("criteria"):
        cfg["criteria"] = spec["criteria"]
    out = {"name": name, "root": root, "spec": spec, "expected": exp, "yaml": cfg}
    if tmp_path is not None:
        import yaml
        cfg = dict(cfg, out=str(tmp_path / "out"))
        path = str(tmp_path / "config.yml")
        with open(path, "w") as f:
            yaml.safe_dump(cfg, f, sort_keys=False)
        out["config_path"] = path
    return out


def fnum(v):
    """Inverse of make_golden.f: non-finite floats were stored as repr strings."""
    if isinstance(v, str):
        return float(v)
    return v


@pytest.fixture(params=CASES)
def golden_case(request, tmp_path):
    return load_case(request.param, tmp_path)
