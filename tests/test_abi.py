"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol
include/moondream_b200.h declares, and the ctypes table matches the header."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "moondream_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(md_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from moondream_b200 import _native as N

    lib = ctypes.CDLL(N.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_table_matches_header():
    from moondream_b200 import _native as N

    assert N.exported_symbols() == _header_functions()
    lib = N.lib()
    assert lib.md_abi_version() == 3
    assert lib.md_linear_small_batch_splits(2048, 8192) >= 1


def test_struct_layouts():
    from moondream_b200 import _native as N

    assert ctypes.sizeof(N.md_dims) == 24 * 4
    assert ctypes.sizeof(N.md_kv) == 40          # ptr, int(+pad), ptr, int, int, int(+pad)


def test_errors_are_reported_not_crashed():
    from moondream_b200 import _native as N

    lib = N.lib()
    rc = lib.md_linear_bf16(None, 0, None, 0, 1, 8, 8, 0, None, None, 0, 0, None, 0, 0, 0, 0, None)
    assert rc != 0 and b"null" in lib.md_last_error()
    rc = lib.md_model_num_weights(None)
    assert rc == -1


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from moondream_b200 import config as C, synth
    from moondream_b200.engine import Engine
    from moondream_b200._native import NativeError

    cfg = C.tiny()
    with pytest.raises(NativeError):
        Engine(cfg, synth.synthetic_state_dict(cfg, 0))
