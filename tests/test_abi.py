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


def test_uvd_gram_plan_geometry_fits_the_sm_for_every_rank():
    """psgd_uvd_plan_info (no GPU needed): the update Gram sweep's pipeline geometry per rank stays launchable on sm_100
    (<= 1024 threads, >= 4 stages of whole-warp row tiles inside the 160 KB ring + static tables under 227 KB)."""
    import ctypes as C
    from psgd_tf_b200 import _lib
    lib = _lib.load_library()
    for r in range(1, 17):
        tile, stages, threads, roles = (C.c_int() for _ in range(4))
        assert lib.psgd_uvd_plan_info(r, C.byref(tile), C.byref(stages), C.byref(threads), C.byref(roles)) == 0
        assert tile.value % 32 == 0 and tile.value >= 32
        assert 4 <= stages.value <= 8
        assert threads.value % 32 == 0 and 64 <= threads.value <= 1024
        assert 1 <= roles.value <= 9
        ring = stages.value * tile.value * 4 * (2 * r + 3)
        assert ring <= 160 * 1024
        table = (2 * r + 2) ** 2 * 4 * 8                      # per-warp-in-role tables (static shared memory), upper bound
        assert ring + table <= 227 * 1024
    assert lib.psgd_uvd_plan_info(17, C.byref(tile), C.byref(stages), C.byref(threads), C.byref(roles)) != 0
