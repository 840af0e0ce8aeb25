"""The C-ABI library loads and exports every symbol include/ssr_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssrb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ("ssrb_lm_create", "ssrb_lm_begin", "ssrb_lm_decode", "ssrb_codec_encode", "ssrb_codec_decode",
                 "ssrb_codec_wmdecode", "ssrb_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_binding_matches_header(built_lib):
    from ssr_speech_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()
    lib = _lib.load()
    assert lib.ssrb_version() >= 100
    assert lib.ssrb_launch_count() == 0 or lib.ssrb_launch_count() > 0


def test_struct_layouts_match_header():
    """ctypes mirrors must have the same field count/order as the C structs (all-int layouts)."""
    from ssr_speech_b200 import _lib
    src = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} ssrb_lm_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"int\s+([^;]+);", body) for n in decl.split(",")]
    assert names == [f[0] for f in _lib.LMConfig._fields_]
    assert ctypes.sizeof(_lib.LMConfig) == 4 * len(names)


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    from ssr_speech_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
