"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/oasr.h declares."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from omnilingual_asr import _native

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "oasr.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"OASR_API\s+[\w\s\*]+?\b(oasr_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    if not _native.lib_path().exists():
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(str(_native.lib_path()))


def test_header_declares_what_python_binds():
    assert set(declared_symbols()) == set(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), name


def test_no_compute_call_needed_for_metadata(built_lib):
    built_lib.oasr_version.restype = ctypes.c_char_p
    assert b"sm_100a" in built_lib.oasr_version()
    cfg = _native.OasrConfig()
    cfg.n_fe_layers = 7
    for i, (k, s) in enumerate([(10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2)]):
        cfg.fe_kernel[i], cfg.fe_stride[i] = k, s
    built_lib.oasr_feature_length.restype = ctypes.c_int32
    built_lib.oasr_feature_length.argtypes = [ctypes.POINTER(_native.OasrConfig), ctypes.c_int64]
    assert built_lib.oasr_feature_length(ctypes.byref(cfg), 480000) == 1499
    assert built_lib.oasr_feature_length(ctypes.byref(cfg), 399) == 0


def test_sass_is_blackwell_native():
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped binary (UTC*MMA, LDTM, UTMALDG)."""
    out = subprocess.run(["cuobjdump", "-sass", str(_native.lib_path())], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out.stdout, mnemonic
    assert "HMMA.16816" not in out.stdout   # no legacy mma.sync path


def test_product_path_fails_loudly_without_library(monkeypatch, tmp_path):
    monkeypatch.setenv("OASR_LIB", str(tmp_path / "missing.so"))
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.load()


def test_product_never_imports_the_oracle():
    for p in (ROOT / "omnilingual-asr_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
