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


# ---- host-side planning entry points: pure functions of the geometry, callable without a GPU -------------------------------------
# (Cin, hid, Cout, patch) of the shipped inverted-residual levels and their arranged row lengths: B1 [ceil(Cin/8)][hid][8] |
# W2T [9][hid] padded to 16 bytes | B2 [ceil(hid/8)][Cout][8]
IR_ROWS = {(34, 68, 19, 16): 5 * 68 * 8 + 616 + 9 * 19 * 8, (26, 52, 19, 16): 4 * 52 * 8 + 472 + 7 * 19 * 8,
           (22, 44, 12, 16): 3 * 44 * 8 + 400 + 6 * 12 * 8, (24, 48, 16, 8): 3 * 48 * 8 + 432 + 6 * 16 * 8,
           (14, 28, 8, 8): 2 * 28 * 8 + 256 + 4 * 8 * 8}


def test_arranged_row_layout_and_supported_shapes(lib):
    for (cin, hid, cout, ps), row in IR_ROWS.items():
        assert lib.hsb_ir_arranged_row_elems(cin, hid, cout) == row, (cin, hid, cout)
        assert row % 8 == 0                                       # a row is a whole number of 16-byte units
        assert lib.hsb_patch_ir_arranged_supported(cin, hid, cout, ps) == 1
        assert lib.hsb_patch_ir_arranged_supported(cin, hid, cout, 24 - ps) == 0      # the other patch size is not instantiated
    assert lib.hsb_patch_ir_arranged_supported(20, 40, 8, 16) == 0
    assert lib.hsb_ir_arranged_row_elems(0, 4, 4) == -1


# (sig_index, sig_ch, head outputs, groups, hp_offset, Cin, hid, Cout): HyperSeg-M levels 4 / 3, unify's shared head (S-Cityscapes)
ARRANGED_PLANS = [(0, 320, 4216, 4, 0, 34, 68, 19), (0, 192, 2352, 16, 0, 24, 48, 16), (768, 512, 3680, 16, 868, 26, 52, 19),
                  (768, 512, 3680, 16, 0, 14, 28, 8), (0, 128, 1896, 8, 0, 22, 44, 12), (4, 192, 2352, 16, 0, 24, 48, 16)]


@pytest.mark.parametrize("geom", ARRANGED_PLANS)
def test_arranged_head_plan_is_consistent(lib, geom):
    """hsb_head_arranged_plan: enough 128-column tiles to cover the arranged row, K a multiple of 16 that holds at least one
    group's signal channels, packed size = tiles x 128 x K; geometry errors are reported, not planned around."""
    import ctypes
    si, sc, och, g, off, cin, hid, cout = geom
    elems, items, kmax = ctypes.c_int64(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = lib.hsb_head_arranged_plan(si, sc, och, g, off, cin, hid, cout, ctypes.byref(elems), ctypes.byref(items), ctypes.byref(kmax))
    assert rc == 0, lib.hsb_last_error()
    row = lib.hsb_ir_arranged_row_elems(cin, hid, cout)
    assert items.value * 128 >= row and items.value <= 2 * (-(-row // 128))
    assert kmax.value % 16 == 0 and kmax.value >= sc // g and kmax.value <= 256
    assert elems.value == items.value * 128 * kmax.value
    # the block's hyper-parameters must fit the head: one output too few is an error
    hp = cin * hid + 9 * hid + hid * cout
    short = (off + hp - 1) // g * g
    assert lib.hsb_head_arranged_plan(si, sc, short, g, off, cin, hid, cout, None, None, None) != 0
    assert lib.hsb_head_arranged_plan(si, sc + 1, och, g, off, cin, hid, cout, None, None, None) != 0      # sig_ch % groups


@pytest.mark.parametrize("geom", [(416, 5248, 32), (224, 3008, 16), (128, 704, 8), (192, 2352, 16), (320, 4216, 4), (576, 4160, 32),
                                  (2048, 4096, 2)])
def test_packed_head_size(lib, geom):
    """hsb_head_packed_elems: never smaller than the weights themselves padded to 16 signal channels per group; the last
    geometry (1024 signal channels per group: the slice cannot stay resident in one CTA) takes the legacy layout."""
    sc, och, g = geom
    n = lib.hsb_head_packed_elems(sc, och, g)
    spg = sc // g
    assert n >= och * (-(-spg // 16) * 16)
    assert n % 8 == 0
    assert lib.hsb_head_packed_elems(sc + 1, och, g) == -1
