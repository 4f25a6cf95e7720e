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


def test_every_entry_point_reports_null_arguments():
    """Error behaviour of the boundary: called with null pointers and zero sizes, every entry point that
    returns a status must come back non-zero with md_last_error() set — before touching CUDA, so this runs without a
    GPU.  In a subprocess: a crash would otherwise take the test session with it."""
    import subprocess
    import sys

    code = r'''
import ctypes, sys
from moondream_b200 import _native as N
lib = N.lib()
void = [n for n, (res, _) in N._SIGNATURES.items() if res is None]
info = {"md_last_error", "md_abi_version", "md_launch_count", "md_linear_small_batch_splits",
        "md_linear_small_batch_workspace_bytes", "md_debug_timeline"}
bad = []
for name, (res, args) in N._SIGNATURES.items():
    if name in void or name in info:
        continue
    vals = [0 if a in (ctypes.c_int, ctypes.c_longlong, ctypes.c_uint) else 0.0 if a is ctypes.c_float else None for a in args]
    rc = getattr(lib, name)(*vals)
    if rc == 0:
        bad.append(name)
    elif res is ctypes.c_int and rc > 0 and not lib.md_last_error():
        bad.append(name + " (no message)")
print("BAD", bad)
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    assert "BAD []" in r.stdout, r.stdout
