"""The C-ABI library loads without a GPU and exports every symbol include/duett_b200.h declares; argument errors are
reported through dx_last_error (no compute calls here)."""
import ctypes as C

import pytest


def test_build_and_exports():
    import __graft_entry__ as g
    g.build()
    from multimodal_edema_prediction_b200 import _decl, _lib
    lib = _lib.lib()
    declared = g.declared_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(lib, s), s
    # every bound signature is declared in the header, and vice versa (dx_last_error / dx_version / dx_device_ok /
    # dx_gemm* are bound in _lib.py)
    bound = set(_decl.SIGNATURES) | {"dx_last_error", "dx_version", "dx_device_ok", "dx_gemm", "dx_gemm_tc_debug",
                                         "dx_gemm_reserve_sms"}
    assert bound == set(declared), (bound ^ set(declared))
    assert lib.dx_version() >= 100
    prev = lib.dx_gemm_reserve_sms(9)                  # host-only setter: rounds down to an even count, returns the previous value
    assert lib.dx_gemm_reserve_sms(prev) == 8


def test_argument_errors_do_not_need_a_gpu():
    from multimodal_edema_prediction_b200 import _lib
    lib = _lib.lib()
    d = _lib.GemmDesc()          # all zeros: empty problem
    rc = lib.dx_gemm(C.byref(d), None)
    assert rc == -1 and b"empty problem" in lib.dx_last_error()
    with pytest.raises(_lib.DxError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from multimodal_edema_prediction_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DxError, match="no CPU fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from multimodal_edema_prediction_b200.duett.duett import Model
    m = Model(3, 5, 1, d_embedding=8, masked_transform_timesteps=4, max_len=4, pretrain=False)
    x = (torch.zeros(2, 3), torch.zeros(2, 4, 11), torch.zeros(2, 4), [4, 4])
    with pytest.raises(Exception, match="no CPU path"):
        m.encode(x)
