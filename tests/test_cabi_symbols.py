"""The C-ABI shared library loads without a GPU and exports every function include/*.h declares;
the ctypes prototype table covers exactly that set.  No compute call is made here."""
import ctypes
import os
import re

from conftest import REPO


def _declared():
    names = set()
    inc = os.path.join(REPO, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(sige_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(cabi_lib):
    lib = ctypes.CDLL(cabi_lib)
    declared = _declared()
    assert len(declared) >= 14
    for name in sorted(declared):
        assert hasattr(lib, name), "libsige_b200.so does not export %s" % name


def test_prototype_table_matches_header(cabi_lib):
    from sige_b200 import _cabi

    assert set(_cabi.PROTOTYPES) == _declared()
    lib = _cabi.lib()
    assert lib.sige_abi_version() == 1
    assert lib.sige_built_arch() == b"sm_100a"
    assert lib.sige_activation_from_name(b"identity") == 0
    assert lib.sige_activation_from_name(b"swish") == 1
    assert lib.sige_activation_from_name(b"gelu") == -1     # the reference has UB here (common.cpp:22)
    assert lib.sige_reduce_mask_capacity(256, 256, 6, 6, 4, 4, 1, 1) == 65 * 65


def test_argument_errors_are_reported_not_crashes(cabi_lib):
    """Bad arguments return non-zero with a message before any CUDA call (safe without a GPU)."""
    from sige_b200 import _cabi

    lib = _cabi.lib()
    rc = lib.sige_gather(None, 0, 7, 1, 1, 4, 4, 2, 2, None, 1, None, None, 0, 0, None, None)
    assert rc != 0 and b"layout" in lib.sige_last_error()
    assert lib.sige_gather(None, 0, 0, 1, 1, 4, 4, 2, 2, None, 0, None, None, 0, 0, None, None) == 0   # N == 0: no-op
    rc = lib.sige_tile_conv(None, None)
    assert rc != 0 and b"null" in lib.sige_last_error()


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from sige_b200 import _cabi

    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _cabi.lib()
    except _cabi.SigeLibraryMissing as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("expected SigeLibraryMissing")


def test_attention_shape_predicate_is_pure_host_logic(cabi_lib):
    """sige_attention_tokens_supported() never touches the device: the shapes the fused attention core takes
    (64 keys per CTA, clusters of 1 / 2 / 4 key slices; 256 or 512 channels; 16-bit storage)."""
    h = ctypes.CDLL(cabi_lib)
    f = h.sige_attention_tokens_supported
    f.restype, f.argtypes = ctypes.c_int, [ctypes.c_int] * 3
    F32, F16, BF16 = 0, 1, 2
    assert f(256, 512, F16) == 1 and f(64, 512, BF16) == 1 and f(128, 256, F16) == 1
    assert f(256, 512, F32) == 0      # fp32 goes through the library path
    assert f(100, 512, F16) == 0 and f(512, 512, F16) == 0 and f(256, 192, F16) == 0 and f(0, 512, F16) == 0
