"""CPU checks of the drop-in boundary: both libraries load and export every symbol of include/meepo.h."""
import os
import re

import pytest

from meepoembedding_b200 import _capi as capi

from conftest import ORACLE_SO, PRODUCT_SO, ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "meepo.h")).read()
    return sorted(set(re.findall(r"MEEPO_API\s+[\w\s\*]+?\b(meepo_\w+)\s*\(", text)))


def test_binding_covers_header():
    assert header_symbols() == sorted(capi.SIGNATURES)


@pytest.mark.parametrize("path", [PRODUCT_SO, ORACLE_SO])
def test_library_exports_every_symbol(path, oracle_lib):
    if path == PRODUCT_SO and not os.path.exists(path):
        import __graft_entry__

        __graft_entry__.build()
    lib = capi.load_library(path)  # binds every name in SIGNATURES or raises
    assert lib.abi_version() == capi.ABI_VERSION
    assert lib.backend_name == ("cuda-sm_100a" if path == PRODUCT_SO else "oracle-cpu")
    assert lib.last_error() is not None
    assert lib.owner(12345, 8) < 8


def test_owner_agrees_between_libraries(oracle_lib):
    cuda = capi.load_library(PRODUCT_SO)
    for k in [0, 1, 2**63, 0xFFFFFFFFFFFFFFFD, 0x9E3779B97F4A7C15]:
        for g in (1, 2, 3, 8):
            assert cuda.owner(k, g) == oracle_lib.owner(k, g)


def test_product_path_does_not_touch_the_oracle():
    """No file of the product package may name the oracle (ROOT/oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "meepoembedding_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dp, f)).read()
                assert "libmeepo_oracle" not in src and "oracle/" not in src, f
