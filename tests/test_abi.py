"""CPU: the C-ABI library loads and exports every symbol include/super_b200.h declares."""
import ctypes
import os

from super_b200 import lib, build


def test_header_parses_and_lists_entry_points():
    sigs = lib.parse_header()
    for name in ("sb_knn", "sb_knn_weights", "sb_warp_update", "sb_data_term_jtj", "sb_data_term_loss",
                 "sb_data_term_rows", "sb_reg_terms", "sb_lm_begin", "sb_lm_damp", "sb_lm_step", "sb_lm_decide"):
        assert name in sigs, name
    assert sigs["sb_knn"] == "pippiiippp"


def test_library_exports_every_declared_symbol():
    path = build.build()                      # no-op when up to date; nvcc cross-compiles without a GPU
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    for name in lib.parse_header():
        assert hasattr(dll, name), f"{name} declared in include/super_b200.h but not exported"
    assert dll.sb_version() >= 100
    assert dll.sb_lm_state_bytes() > 0       # host-only queries are safe without a GPU
    assert dll.sb_data_loss_blocks(300000) == 592


def test_no_silent_fallback_when_library_missing(monkeypatch):
    import pytest
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libsuper_b200.so")
    with pytest.raises(lib.SuperB200Error):
        lib.load()
