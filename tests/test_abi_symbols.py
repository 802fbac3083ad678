"""The C-ABI library loads and exports exactly the symbols include/zkb200.h declares (no compute calls, no GPU)."""
import ctypes as C
import os
import re

from zktls_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_shared_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.SO_PATH), "libzkb200.so missing: run python __graft_entry__.py"
    L = C.CDLL(_lib.SO_PATH)
    for name in header_symbols():
        assert hasattr(L, name), name
    L.zkb_version.restype = C.c_char_p
    assert L.zkb_version().startswith(b"zkb200")


def test_init_without_gpu_is_an_error_string_not_a_crash():
    import torch
    if torch.cuda.is_available():
        return
    L = _lib.lib()
    ctx = C.c_void_p()
    err = L.zkb_init(C.c_int(0), C.byref(ctx))
    assert err, "zkb_init must fail without a CUDA device (no CPU fallback)"
    L.zkb_free_error(C.c_void_p(err))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zktls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "#include \"../../oracle" not in src, f


def test_generated_rust_bindings_are_current():
    """integration/rust/zkb200-sys/src/lib.rs is generated from the header (tools/gen_rust_ffi.py) and must not drift."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    src, names = mod.generate()
    assert sorted(names) == header_symbols()
    assert open(os.path.join(ROOT, "integration", "rust", "zkb200-sys", "src", "lib.rs")).read() == src, "run python tools/gen_rust_ffi.py"
