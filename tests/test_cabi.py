"""CPU: the C-ABI library builds, loads and exports every symbol include/nnmpc.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nnmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnmpc_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("nnmpc_qp_create", "nnmpc_qp_solve", "nnmpc_qp_solve_host", "nnmpc_ts_solve", "nnmpc_sim_run",
              "nnmpc_sim_run_host", "nnmpc_mlp_forward", "nnmpc_mlp_forward_host", "nnmpc_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for s in _declared_symbols():
        assert hasattr(handle, s), f"{s} declared in include/nnmpc.h but not exported"


def test_ctypes_table_covers_header(built_lib):
    from industrial_nnmpc_2021_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    L = _lib.lib()
    assert L.nnmpc_version() >= 100
    assert L.nnmpc_launch_count() >= 0


def test_create_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: on a box without a CUDA device the product raises instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np
    from industrial_nnmpc_2021_b200 import _lib
    from industrial_nnmpc_2021_b200.linearMPC import DenseQPRegulator
    A = np.array([[0.9]]); B = np.array([[1.0], ])
    with pytest.raises(_lib.NnmpcError):
        DenseQPRegulator(A=np.array([[0.9, 0.0], [0.0, 0.0]]), B=np.array([[1.0], [1.0]]), Q=np.eye(2), R=np.eye(1),
                         M=np.zeros((2, 1)), N=4, ulb=-np.ones((1, 1)), uub=np.ones((1, 1)))
    # raw C call reports the reason
    L = _lib.lib()
    h = ctypes.c_void_p()
    P = np.eye(2)
    rc = L.nnmpc_qp_create(ctypes.byref(h), 2, 2, 1, 2, _lib.hptr(P), _lib.hptr(P), _lib.hptr(P), _lib.hptr(P),
                           _lib.hptr(P), 1.6, 0)
    assert rc < 0 and b"CUDA" in L.nnmpc_last_error()
