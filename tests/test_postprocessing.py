"""``xpore postprocessing`` against outputs of the unmodified reference filter
(fixtures from tests/golden/make_golden_post.py)."""
import os
import shutil

import pytest

from xpore_b200 import cli
from xpore_b200.postprocessing import run_postprocessing

POST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post")
CASES = sorted(f[:-len(".diffmod.table")] for f in os.listdir(POST) if f.endswith(".diffmod.table"))


@pytest.mark.parametrize("name", CASES)
def test_matches_reference_output(name, tmp_path):
    out = run_postprocessing(os.path.join(POST, name + ".diffmod.table"), str(tmp_path))
    assert os.path.basename(out) == "majority_direction_kmer_diffmod.table"
    with open(out) as f, open(os.path.join(POST, name + ".expected.table")) as g:
        assert f.read() == g.read()


def test_cli_subcommand(tmp_path):
    shutil.copy(os.path.join(POST, "mixed.diffmod.table"), tmp_path / "diffmod.table")
    cli.main(["xpore", "postprocessing", "--diffmod_dir", str(tmp_path)])
    with open(tmp_path / "majority_direction_kmer_diffmod.table") as f, open(os.path.join(POST, "mixed.expected.table")) as g:
        assert f.read() == g.read()


def test_idempotent(tmp_path):
    first = run_postprocessing(os.path.join(POST, "mixed.diffmod.table"), str(tmp_path))
    d2 = tmp_path / "again"
    d2.mkdir()
    second = run_postprocessing(first, str(d2))
    with open(first) as f, open(second) as g:
        assert f.read() == g.read()
