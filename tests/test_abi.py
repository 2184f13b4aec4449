"""The C-ABI library loads and exports every symbol include/psgd_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from psgd_tf_b200 import build, _lib
    build.build()
    return _lib.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "psgd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psgd_[a-z0-9_]+)\s*\(", text)) - {"psgd_allreduce_fn"})


def test_header_symbols_are_exported_and_typed(lib):
    from psgd_tf_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/psgd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.psgd_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.psgd_create(0, None, C.byref(h))
    assert rc == 6 and b"no CPU path" in lib.psgd_last_error()
    import psgd_tf_b200 as psgd
    x = torch.ones(4, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        psgd.precond_grad_kron(x, x, x)
    with pytest.raises(RuntimeError, match="no CPU"):
        psgd.get_context()


def test_layer_struct_matches_header_layout():
    from psgd_tf_b200 import _lib
    # 2 x int32, 2 x int64, 8 pointers
    assert C.sizeof(_lib.KronLayer) == 8 + 16 + 8 * 8


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under psgd_tf_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "psgd_tf_b200")
    pat = re.compile(r"^\s*(from|import)\s+.*oracle|psgd_oracle|oracle[/.]_ref", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f"{f} references the oracle"
