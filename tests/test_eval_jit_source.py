"""Host-only checks of the eval_check code generator (k_eval_jit.cu): the CUDA source it emits for a circuit compiles for
sm_100a with NVRTC -- no device needed -- in both of its forms, and the staged form really is the bulk-copy pipeline."""
import ctypes as C
import os

import numpy as np
import pytest

from zktls_b200 import circuit
from zktls_b200._lib import check, lib

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
MID = dict(accum_cols=7, code_cols=5, data_cols=33, mix_size=20, out_size=32)


def _source(blob):
    L = lib()
    bp = blob.ctypes.data_as(C.POINTER(C.c_uint32))
    need = C.c_size_t(0)
    check(L.zkb_eval_check_source(bp, C.c_size_t(blob.size), None, C.c_size_t(0), C.byref(need)))
    buf = C.create_string_buffer(need.value + 1)
    check(L.zkb_eval_check_source(bp, C.c_size_t(blob.size), buf, C.c_size_t(need.value + 1), C.byref(need)))
    return buf.value.decode()


@pytest.fixture
def cache_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("ZKB_CACHE_DIR", str(tmp_path))      # do not touch the cubins shipped in zktls_b200/_jitcache
    return tmp_path


@pytest.mark.parametrize("shape", [SMALL, MID])
def test_staged_source_is_a_bulk_copy_pipeline_and_compiles(shape, cache_dir, monkeypatch):
    monkeypatch.setenv("ZKB_EC_STAGED", "1")
    blob = circuit.syn_circuit(**shape).blob()
    src = _source(blob)
    assert "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes" in src
    assert "mbarrier.try_wait.parity" in src and "extern __shared__" in src
    # every streamed column of a block is copied before the block is waited for: as many copy_col calls as slots named
    assert src.count("copy_col(") >= shape["accum_cols"] + shape["data_cols"]
    try:
        check(lib().zkb_eval_check_precompile(blob.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(blob.size)))
    except Exception as e:      # noqa: BLE001
        if "libnvrtc" in str(e):
            pytest.skip("libnvrtc not available")
        raise
    assert [f for f in os.listdir(cache_dir) if f.endswith(".zkbj")], "no cubin written"


def test_register_form_still_generated_on_request(cache_dir, monkeypatch):
    monkeypatch.setenv("ZKB_EC_STAGED", "0")
    src = _source(circuit.syn_circuit(**SMALL).blob())
    assert "cp.async.bulk" not in src and "__ldg" in src


def test_staged_and_register_forms_evaluate_the_same_program():
    """Same constraint arithmetic in both forms: the lines after the tap loads are identical."""
    blob = circuit.syn_circuit(**SMALL).blob()
    os.environ["ZKB_EC_STAGED"] = "1"
    a = _source(blob)
    os.environ["ZKB_EC_STAGED"] = "0"
    b = _source(blob)
    os.environ.pop("ZKB_EC_STAGED")
    arith = lambda s: [l.strip() for l in s.splitlines() if any(k in l for k in ("wacc(", "= mul(", "= sub(", "= add(", "scale4(", "fin("))
                       and "__device__" not in l]
    assert arith(a) == arith(b) and len(arith(a)) > 10


def test_the_jit_compiler_does_not_depend_on_import_order(tmp_path):
    """A process that imported torch first has torch's bundled libnvrtc mapped under the SONAME libnvrtc.so.12; the loader must still pick
    the same (newest) NVRTC as a process that did not -- the compact eval_check form ran 8 % slower when it silently got the older one
    (profiles/r2_y_nvrtc_version.txt), and the on-disk cubin cache is keyed by the NVRTC version, so the shipped cubins were missed too."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    body = ("import ctypes as C, numpy as np\n"
            "from zktls_b200 import circuit\n"
            "from zktls_b200._lib import check, lib\n"
            "b = circuit.syn_circuit(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4).blob()\n"
            "check(lib().zkb_eval_check_precompile(b.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(b.size)))\n")
    seen = {}
    for tag, prefix in (("plain", ""), ("torch first", "import torch\n")):
        env = dict(os.environ, ZKB_EC_VERBOSE="1", ZKB_CACHE_DIR=str(tmp_path / tag.replace(" ", "_")), PYTHONPATH=root)
        env.pop("ZKB_NVRTC_LIB", None)
        os.makedirs(env["ZKB_CACHE_DIR"], mode=0o700)
        r = subprocess.run([sys.executable, "-c", prefix + body], env=env, capture_output=True, text=True, timeout=600)
        if "not loadable" in r.stderr:
            pytest.skip("no NVRTC in this environment")
        assert r.returncode == 0, r.stderr[-2000:]
        m = re.search(r"zkb200: NVRTC (\d+)\.(\d+) from (\S+)", r.stderr)
        assert m, r.stderr[-2000:]
        seen[tag] = (int(m.group(1)), int(m.group(2)), m.group(3))
    assert seen["plain"] == seen["torch first"], seen
