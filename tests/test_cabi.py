"""C-ABI surface: the shared library loads, exports every symbol include/ggnn_b200.h declares, and its
host-only entry points (graph configuration, launch-shape derivation, argument validation) agree with
the oracle.  No kernel is launched here (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ggnn_b200 import _lib
from oracle import pyoracle as O
from tests.test_oracle_golden import KAT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ggnn_b200.h")).read()
    declared = set(re.findall(r"\b(ggnn_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/ggnn_b200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert b"sm_100a" in _lib.lib().ggnn_b200_version()


@pytest.mark.parametrize("row", KAT, ids=[f"N{r[0]}_K{r[1]}" for r in KAT])
def test_graph_config_matches_reference_known_answers(row):
    N, K, KF, S, G, S0, S0_off, SG, SG_off, Bs, Ns, N_all, ST_all, blob = row
    c = _lib.graph_config(N, 128, K)
    o = O.graph_config(N, 128, K)
    for f, _t in c._fields_:
        a, b = getattr(c, f), getattr(o, f)
        assert (list(a) == list(b)) if not isinstance(a, int) else (a == b), f
    assert _lib.lib().ggnn_b200_graph_blob_bytes(C.byref(c)) == blob
    off = _lib.graph_offsets(c)
    assert off.translation == N_all * K * 4 and off.selection == off.translation + ST_all * 4
    assert off.nn1_stats == off.translation + ST_all * 8 and off.total == blob


@pytest.mark.parametrize("D,K,it", [(128, 10, 400), (96, 10, 400), (128, 10, 200), (128, 100, 400), (128, 10, 1000),
                                    (256, 10, 400), (960, 1, 64), (4, 40, 2000)])
def test_query_shape_matches_oracle(D, K, it):
    s = _lib.query_shape(D, K, it)
    assert (s.cache_size, s.sorted_size, s.block_dim_x) == O.query_launch_params(D, K, it)


def test_invalid_arguments_are_rejected_like_the_reference_checks():
    with pytest.raises(ValueError):
        _lib.graph_config(1000, 0, 24)        # D >= 1 (ggnn.cuh:48)
    with pytest.raises(ValueError):
        _lib.graph_config(1000, 128, 1)       # KBuild >= 2 (ggnn.cuh:51)
    with pytest.raises(ValueError):
        _lib.graph_config(1000, 128, 513)
    with pytest.raises(ValueError):
        _lib.query_shape(128, 6001, 400)      # query_kernels.cu:66-75
    with pytest.raises(ValueError):
        _lib.query_shape(128, 10, 8193)       # :105
    with pytest.raises(ValueError):
        _lib.query_shape(5000, 10, 400)
    lib = _lib.lib()
    p = _lib.QueryParams()
    assert lib.ggnn_b200_query(C.byref(p), 10, None) == _lib.ERR_INVALID          # null pointers
    assert b"null" in lib.ggnn_b200_last_error()
    b = _lib.BfQueryParams()
    assert lib.ggnn_b200_bf_query(C.byref(b), 10, None) == _lib.ERR_INVALID
    assert lib.ggnn_b200_merge_topk(None, None, 1, 0, 0, 1, 1, 1, 0, None, None, None) == _lib.ERR_INVALID
    cfg = _lib.graph_config(10000, 128, 24)
    assert lib.ggnn_b200_merge(C.byref(cfg), None, 0, 0.5, 1, 1, None, None, None, None) == _lib.ERR_INVALID
    assert lib.ggnn_b200_build_scratch_bytes(C.byref(cfg)) > 10000 * (4 + 24 * 4 + 12 * 4)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "ggnn_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "ggnn_oracle" not in src and "oracle/" not in src, f
