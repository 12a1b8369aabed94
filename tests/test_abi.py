"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from brdf_nerf_b200 import build
    return build.build(verbose=False)


def _declared():
    text = open(os.path.join(ROOT, "include", "brdfnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header(lib_path):
    from brdf_nerf_b200 import _lib
    assert set(_lib.exported_symbols()) == set(_declared())
    lib = _lib.load()
    assert lib.bn_abi_version() == 1
    assert lib.bn_last_error() is not None


def test_errors_without_gpu(lib_path):
    """Argument validation happens before any CUDA call: bad calls return BN_ERR_ARG with a message."""
    from brdf_nerf_b200 import _lib
    lib = _lib.load()
    rc = lib.bn_sample_guided(None, None, None, None, None, None, None, None, 3.0, None, None, 1, None, None, None, None,
                              4, 64, 64, None)
    assert rc == -1 and b"null pointer" in lib.bn_last_error()
    rc = lib.bn_composite_sigma(None, None, None, 0.0, None, None, None, None, None, 4, 64, None)
    assert rc == -1


def test_sm100a_sass_has_tcgen05_and_tma(lib_path):
    """The shipped binary contains Blackwell tensor-core / TMA instructions (UTC*MMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS"


def test_binding_arity_matches_header():
    """Every ctypes signature in brdf_nerf_b200/_lib.py has as many arguments as the header's prototype (ABI drift guard)."""
    from brdf_nerf_b200 import _lib
    text = open(os.path.join(ROOT, "include", "brdfnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(bn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    checked = 0
    for name, (_, argtypes) in _lib._SIGS.items():
        assert name in protos, f"{name} is bound but not declared in the header"
        params = protos[name].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), f"{name}: header has {n} parameters, the binding {len(argtypes)}"
        checked += 1
    assert checked >= 40
