"""C-ABI boundary checks that need no GPU: the library loads, exports every symbol the header declares,
the Python prototypes cover them, and the product path refuses CPU tensors / a missing library loudly."""
import ctypes
import os
import re

import pytest
import torch

from open3dsot_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(o3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 17
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/o3d_b200.h but not exported"


def test_python_prototypes_match_header():
    assert sorted(_lib.PROTOTYPES) == _header_symbols()


def test_version_and_thread_rule():
    L = _lib.lib()
    assert L.o3d_version() == 100
    assert [L.o3d_opt_threads(n) for n in (1, 100, 512, 1024)] == [1, 64, 512, 512]


def test_cpu_tensors_are_rejected_like_upstream():
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.furthest_point_sampling(torch.zeros(1, 8, 3), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ball_query(torch.zeros(1, 2, 3), torch.zeros(1, 8, 3), 0.3, 4)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libo3d_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_argument_errors_return_status_not_crash():
    L = _lib.lib()
    assert L.o3d_fps(None, 1, 8, 4, None, None) < 0
    assert b"null" in L.o3d_last_error()
    assert L.o3d_ballquery_group(1, 1, 1, 1, 8, 4, 3, 0.3, 4, 0, None, 1, None) < 0   # C % 4 != 0
