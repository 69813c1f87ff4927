"""The C ABI: libhsb200.so loads, exports every symbol include/hsb200.h declares, validates arguments and
reports errors -- all without touching a GPU (no compute call is made here)."""
import ctypes
import os
import re

import pytest

from hyperseg_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "hsb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        from hyperseg_b200.build import build
        build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    syms = declared_symbols()
    assert syms == sorted(_lib.SIGNATURES), "ctypes table and header drifted apart"
    assert len(syms) >= 9


def _prototypes():
    """name -> list of parameter kinds ('ptr' / 'i64' / 'int') parsed from the header."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int64_t|int|const\s+char\s*\*)\s+(hsb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = [p.strip() for p in m.group(2).split(",") if p.strip() and p.strip() != "void"]
        out[m.group(1)] = ["ptr" if "*" in p else ("i64" if "int64_t" in p else "int") for p in params]
    return out


def test_ctypes_signatures_match_the_header_prototypes(lib):
    """Arity and pointer / int64 / int kind of every parameter: a drifted signature would otherwise only fail (or
    silently corrupt arguments) at call time on a GPU."""
    protos = _prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_int64: "i64"}
    for name, argtypes in _lib.SIGNATURES.items():
        got = [kinds.get(a, "ptr") for a in argtypes]          # POINTER(c_int) etc. count as pointers
        assert got == protos[name], f"{name}: ctypes {got} vs header {protos[name]}"


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in hsb200.h but not exported"


def test_version(lib):
    m = re.search(r"#define\s+HSB200_VERSION\s+(\d+)", open(HEADER).read())
    assert _lib.version() == int(m.group(1))


def test_null_pointers_are_rejected_with_message(lib):
    rc = lib.hsb_patch_conv1x1_fwd(None, None, None, None, None, 0, 1, 4, 4, 4, 4, 2, 2, 1, 0, 1, 16, None)
    assert rc == -1
    assert b"null pointer" in lib.hsb_last_error()
    rc = lib.hsb_patch_ir_fwd(None, None, None, None, None, None, None, None, None,
                              1, 4, 8, 4, 4, 4, 2, 2, 0, 0, 1, 100, None)
    assert rc == -1
    rc = lib.hsb_signal2weights_fwd(None, None, None, 1, 0, 8, 8, 8, 1, 2, 2, 32, 4, 1, 0, 1, 8, None)
    assert rc == -1
    with pytest.raises(_lib.HsbError, match="null pointer"):
        _lib.check(rc, "hsb_signal2weights_fwd")


def test_shape_validation_happens_before_any_launch(lib):
    fake = ctypes.c_void_p(0x1000)      # never dereferenced: validation fails first
    # H not divisible by fh
    rc = lib.hsb_patch_conv1x1_fwd(fake, fake, fake, None, None, 0, 1, 4, 4, 5, 4, 2, 2, 1, 0, 1, 16, None)
    assert rc == -1 and b"divisible" in lib.hsb_last_error()
    # residual with Cin != Cout
    rc = lib.hsb_patch_ir_fwd(fake, fake, fake, fake, fake, fake, fake, fake, fake,
                              1, 4, 8, 5, 4, 4, 2, 2, 1, 0, 1, 200, None)
    assert rc == -1 and b"residual" in lib.hsb_last_error()
    # non size-preserving generic conv is refused as unsupported (-2)
    rc = lib.hsb_patch_conv_fwd(fake, fake, fake, None, None, 0, 1, 4, 4, 4, 4, 2, 2, 3, 3, 0, 0, 1, 1, 1, 1, 0, 1, 144, None)
    assert rc == -2
    # bad dtype
    rc = lib.hsb_weights_to_patch_major(fake, fake, 1, 8, 2, 2, 8, 7, None)
    assert rc == -1


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    sm, cc = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.hsb_device_info(ctypes.byref(sm), ctypes.byref(cc)) == -4
    assert b"no CPU path" in lib.hsb_last_error()


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libhsb200.so")
    with pytest.raises(_lib.HsbError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_packed_head_cache_is_tied_to_the_live_tensor_object():
    """The tensor-core head packs its static weights once per (tensor, version); an entry must not be served to
    another tensor that later occupies the same address (key collision) nor survive its tensor."""
    import gc
    import torch
    from hyperseg_b200 import ops
    cache, key = {}, ("ptr", 0, (4, 2), 1, torch.float32)
    a, b = torch.zeros(4, 2), torch.zeros(4, 2)
    assert ops._cache_lookup(cache, key, a) is None
    ops._cache_store(cache, key, a, "packed-a")
    assert ops._cache_lookup(cache, key, a) == "packed-a"
    assert ops._cache_lookup(cache, key, b) is None            # same key, other tensor object: miss
    ops._cache_store(cache, key, b, "packed-b")
    assert ops._cache_lookup(cache, key, b) == "packed-b" and ops._cache_lookup(cache, key, a) is None
    del b
    gc.collect()
    assert cache[key][0]() is None                               # dead owner: the entry can never hit again
    for i in range(70):                                          # live owners are never evicted (a CUDA graph may read them)
        ops._cache_store(cache, ("k", i), a, i)
    assert all(ops._cache_lookup(cache, ("k", i), a) == i for i in range(70))
    assert key not in cache                                      # ... dead ones are pruned once the table grows
