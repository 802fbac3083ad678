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


def test_rust_integration_tree_is_self_consistent():
    """integration/rust is source only (no Rust toolchain here), but it must not dangle: every `crate::x::y` path used under
    guest-prover-b200/src resolves to a module file that defines `y`, every `mod x;` has its file, every zkb200_sys / sys:: call
    names a function the generated bindings declare, and the CLI patch touches the reference files VERDICT r1 listed."""
    src = os.path.join(ROOT, "integration", "rust", "guest-prover-b200", "src")
    files = {f[:-3]: open(os.path.join(src, f)).read() for f in os.listdir(src) if f.endswith(".rs")}
    for name, text in files.items():
        for m in re.findall(r"^\s*pub mod (\w+);", text, flags=re.M):
            assert m in files, f"{name}.rs declares mod {m} but src/{m}.rs is missing"
        for mod, item in re.findall(r"crate::(\w+)::(\w+)", text):
            assert mod in files, f"{name}.rs uses crate::{mod} but src/{mod}.rs is missing"
            assert re.search(rf"\b(fn|struct|trait|const|static|enum|type)\s+{item}\b", files[mod]), f"crate::{mod}::{item} (used in {name}.rs) is not defined"
        assert not re.search(r"crate::[A-Z_]{3,}\b", text), f"{name}.rs refers to a crate-level static that does not exist"
    bindings = open(os.path.join(ROOT, "integration", "rust", "zkb200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (zkb_\w+)\(", bindings)) | {"ffi_wrap"}
    for name, text in files.items():
        for fn in re.findall(r"(?:sys|zkb200_sys)::(zkb_\w+|ffi_wrap)", text):
            assert fn in declared, f"{name}.rs calls {fn}, which zkb200-sys does not declare"
    patch = open(os.path.join(ROOT, "integration", "patches", "0001-bins-zktls-add-b200-backend.patch")).read()
    for touched in ("bins/zktls/src/commands/types.rs", "bins/zktls/src/commands/prove.rs", "bins/zktls/Cargo.toml"):
        assert f"+++ b/{touched}" in patch
    assert "B200," in patch and "b200-backend" in patch and "B200GuestProver" in patch
