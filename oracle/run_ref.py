"""Run the UNMODIFIED reference CLI (``xpore diffmod ...``) from ``oracle/_ref`` - TEST / BENCHMARK INFRASTRUCTURE.

    python oracle/run_ref.py diffmod --config cfg.yml --n_processes 16 [--ids ...]

``oracle/_ref`` is a verbatim copy of the reference's package made by ``oracle/make_ref.sh`` (git-ignored).  The
only additions are the two import shims of SURVEY.md 8c: ``ujson`` -> the standard ``json`` module (used solely
for ``ujson.loads``, ``xpore/scripts/diffmod.py:169``) and an empty ``h5py`` module (used solely by
``--save_models``, ``xpore/diffmod/io.py:93-180``).  Neither package is installed in this image; neither takes part
in the arithmetic.  Only ``bench.py`` executes this file (the reference arm and the CLI comparison).
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "xpore", "scripts", "xpore.py"))


def main(argv):
    if not available():
        sys.exit("oracle/_ref is missing: run oracle/make_ref.sh where /root/reference exists")
    sys.path.insert(0, REF)
    sys.modules.setdefault("ujson", json)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    from xpore.scripts import xpore as ref_cli
    ref_cli.main(["xpore"] + list(argv))


if __name__ == "__main__":
    main(sys.argv[1:])
