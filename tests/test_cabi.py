"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/dad_b200.h declares, and follows its error conventions (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from distill_any_depth_b200 import _lib
    return _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dad_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 24
    h = lib.load()
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/dad_b200.h but not exported"
    assert sorted(lib.PROTOTYPES) == syms, "ctypes prototypes and header out of sync"
    assert h.dad_abi_version() == 1


def test_no_torch_or_libcuda_link_dependency():
    """Pure C ABI: the .so must not pull in torch (or need libcuda at load time)."""
    import subprocess
    from distill_any_depth_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcuda.so" not in out and "not found" not in out, out


def test_error_conventions_without_gpu(lib):
    h = lib.load()
    d = lib.ModelDesc()
    d.embed_dim, d.depth, d.num_heads = 1000, 12, 6  # head_dim != 64
    handle = ctypes.c_void_p()
    rc = h.dad_model_create(ctypes.byref(d), ctypes.byref(handle))
    assert rc == lib.DAD_ERR_UNSUPPORTED
    assert b"head_dim" in h.dad_last_error()
    with pytest.raises(NotImplementedError):
        lib.check(rc, "dad_model_create")
    assert h.dad_model_create(None, ctypes.byref(handle)) == lib.DAD_ERR_INVALID
    with pytest.raises(ValueError):
        lib.check(lib.DAD_ERR_INVALID)
    with pytest.raises(RuntimeError):
        lib.check(lib.DAD_ERR_CUDA)


def test_workspace_sizing_is_consistent(lib):
    """The dry-run arena (no GPU needed) grows with batch / resolution and the fp32 mode needs more."""
    h = lib.load()
    d = lib.ModelDesc()
    d.embed_dim, d.depth, d.num_heads = 384, 12, 6
    d.taps = (ctypes.c_int * 4)(2, 5, 8, 11)
    d.features = 64
    d.out_channels = (ctypes.c_int * 4)(48, 96, 192, 384)
    handle = ctypes.c_void_p()
    assert h.dad_model_create(ctypes.byref(d), ctypes.byref(handle)) == 0
    try:
        w = lambda B, H, W, m: int(h.dad_forward_workspace_bytes(handle, B, H, W, m))
        assert 0 < w(1, 70, 98, 0) < w(2, 70, 98, 0) < w(2, 518, 518, 0)
        assert w(1, 518, 518, 1) > w(1, 518, 518, 0)
        assert abs(w(4, 392, 392, 0) - 4 * w(1, 392, 392, 0)) < 0.02 * w(4, 392, 392, 0)
        assert w(1, 75, 70, 0) == 0 and b"multiples of 14" in h.dad_last_error()
        assert int(h.dad_loss_workspace_bytes(16, 7)) > int(h.dad_loss_workspace_bytes(16, 1)) > 0
    finally:
        h.dad_model_destroy(handle)


def test_training_workspace_sizing_and_argument_checks(lib):
    """dad_train_workspace_bytes is a dry run of the tape + backward-scratch layout (no GPU needed): both engines size,
    the bf16 tape is smaller than the fp32 one, and it exceeds the inference workspace."""
    h = lib.load()
    d = lib.ModelDesc()
    d.embed_dim, d.depth, d.num_heads = 384, 12, 6
    d.taps = (ctypes.c_int * 4)(2, 5, 8, 11)
    d.features = 64
    d.out_channels = (ctypes.c_int * 4)(48, 96, 192, 384)
    handle = ctypes.c_void_p()
    assert h.dad_model_create(ctypes.byref(d), ctypes.byref(handle)) == 0
    try:
        t = lambda B, H, W, m: int(h.dad_train_workspace_bytes(handle, B, H, W, m))
        f = lambda B, H, W, m: int(h.dad_forward_workspace_bytes(handle, B, H, W, m))
        assert t(2, 70, 98, 1) > f(2, 70, 98, 1) > 0
        assert 0 < t(2, 70, 98, 0)
        assert t(4, 392, 392, 0) < t(4, 392, 392, 1)
        assert t(1, 392, 392, 1) < t(2, 392, 392, 1) < t(4, 392, 392, 1)
        assert t(1, 75, 70, 1) == 0 and b"multiples of 14" in h.dad_last_error()
        assert t(1, 70, 70, 2) == 0
        # gradient registration needs a known parameter of the right size; backward without prepared weights is refused
        assert h.dad_model_set_grad(handle, b"pretrained.nope", ctypes.c_void_p(16), 4) == lib.DAD_ERR_INVALID
        assert h.dad_backward(handle, 1, 70, 70, 1, None, None, ctypes.c_void_p(1024), 1 << 40, None) == lib.DAD_ERR_INVALID
        assert b"dad_model_prepare" in h.dad_last_error()
    finally:
        h.dad_model_destroy(handle)
