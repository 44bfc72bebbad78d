"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/xpore_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

from xpore_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    _lib.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "xpore_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xpb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(library):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(library, s), s
    assert sorted(syms) == sorted(_lib.EXPORTS)


def test_record_widths(library):
    from xpore_b200 import fit as xf
    for G, Cn in [(2, 2), (4, 2), (6, 2), (6, 3), (18, 6)]:
        assert library.xpb200_fit_width(G) == xf.fit_width(G)
        assert library.xpb200_stats_width(G, Cn) == xf.stats_width(G, Cn)
    assert library.xpb200_workspace_bytes(1000) >= 4000
    assert b"sm_100a" in library.xpb200_version()


def test_struct_layouts_match_header():
    # field order/sizes of the ctypes mirrors (x86-64 / LP64)
    assert C.sizeof(_lib.XpBatch) == 4 * 4 + 8 + 6 * 8 + 2 * 4 + 8
    assert C.sizeof(_lib.XpMethod) == 2 * 4 + 3 * 8
    assert C.sizeof(_lib.XpResult) == 6 * 8  # status n_iter pval fit stats coef
    assert C.sizeof(_lib.XpLaunchInfo) == 8 * 4


def test_argument_validation_without_gpu(library):
    b = _lib.XpBatch(0, 0, 2, 0, 0, None, None, None, None, None, None, 0, 0, 1.0)
    m = _lib.XpMethod(500, 0, 1e-5, float("nan"), 0.001)
    r = _lib.XpResult(None, None, None, None, None)
    rc = library.xpb200_fit_device(C.byref(b), C.byref(m), C.byref(r), None, 0, None, None)
    assert rc == -1 and b"n_groups" in library.xpb200_last_error()


def test_host_chunk_plan(library):
    """Chunk plans of the host entry points (xp_api.cu:plan_chunks): boundaries strictly ascending from 0 to P, at most 64
    chunks whatever the batch size, chunk-by-chunk sizes growing to 112 MB, chained sizes doubling from 2 MB to 32 MB and
    halving again at the end; small batches are one chunk."""
    import numpy as np
    L = library
    L.xpb200_host_chunk_plan.restype = C.c_int32
    L.xpb200_host_chunk_plan.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]

    def plan(n_reads_per_pos, eb, chained):
        off = np.zeros(len(n_reads_per_pos) + 1, dtype=np.int64)
        np.cumsum(n_reads_per_pos, out=off[1:])
        cuts = np.zeros(80, dtype=np.int32)
        n = L.xpb200_host_chunk_plan(off.ctypes.data, len(n_reads_per_pos), eb, chained, cuts.ctypes.data, len(cuts))
        assert n >= 1, L.xpb200_last_error()
        c = cuts[:n + 1]
        assert c[0] == 0 and c[-1] == len(n_reads_per_pos) and (np.diff(c) > 0).all() and n <= 64
        return off[c[1:]] * eb - off[c[:-1]] * eb  # bytes of reads per chunk

    rng = np.random.default_rng(0)
    c2 = rng.integers(120, 601, size=1_000_000)  # c2 shape: 0.72 GB of u16 reads
    by = plan(c2, 2, 0)
    assert 8 <= len(by) <= 12 and by.max() <= (112 << 20) + 4096 and by[0] < 16 << 20 and (np.diff(by[:-1]) >= -4096).all()
    ch = plan(c2, 2, 1)
    mb = ch / 2**20
    assert 24 <= len(ch) <= 32 and mb[0] < 2.1 and mb.max() < 32.1
    assert (np.diff(mb[:5]) > 0).all() and (np.diff(mb[-3:]) < 0).all() and mb[-1] < 4.1
    assert len(plan(rng.integers(120, 601, size=2000), 2, 1)) == 1  # 1.4 MB: one chunk
    assert len(plan(rng.integers(120, 601, size=2000), 8, 0)) == 1
    big = plan(np.full(3_000_000, 1500), 2, 1)  # 9 GB of reads: still inside the chain table
    assert len(big) <= 64 and big.max() < 9e9 / 40
    off = np.arange(0, 2_000_001 * 360, 360, dtype=np.int64)  # too few entries for the boundaries: an error, not an overrun
    cuts = np.zeros(3, dtype=np.int32)
    assert L.xpb200_host_chunk_plan(off.ctypes.data, 2_000_000, 2, 1, cuts.ctypes.data, 3) == -1
    assert b"cuts holds fewer" in L.xpb200_last_error()


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle or the emulator."""
    pkg = os.path.join(REPO, "xpore_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "xp_emul" not in src, f


def test_torch_custom_op_registered_and_has_no_cpu_path():
    """`torch.ops.xpore_b200.fit` exists after importing xpore_b200.torch_op; shapes follow from the fake kernel; CPU tensors
    are refused (the operator has a CUDA implementation only - no fallback)."""
    import pytest
    import torch
    import xpore_b200.torch_op  # noqa: F401
    from xpore_b200.fit import fit_width, stats_width
    op = torch.ops.xpore_b200.fit
    P, G, Cn = 5, 4, 2
    args = (torch.zeros(40, dtype=torch.float64), torch.arange(0, 48, 8, dtype=torch.int64),
            torch.full((P * G,), 2, dtype=torch.int32), torch.full((P,), 100.0, dtype=torch.float64),
            torch.full((P,), 3.0, dtype=torch.float64), torch.tensor([0, 0, 1, 1], dtype=torch.int32), Cn, 8, 500, 1e-5, 1, 0.1,
            0.001)
    with pytest.raises((NotImplementedError, RuntimeError)):
        op(*args)
    meta = tuple(a.to("meta") if isinstance(a, torch.Tensor) else a for a in args)
    out = op(*meta)
    assert [tuple(t.shape) for t in out] == [(P,), (P,), (P,), (P, fit_width(G)), (P, stats_width(G, Cn)), (1,)]
