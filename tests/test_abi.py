"""The C-ABI library loads and exports every symbol include/spde_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from spdepy_b200 import _lib
    header = open(os.path.join(ROOT, "include", "spde_b200.h")).read()
    names = set(re.findall(r"\b(spde_[a-z0-9_]+)\s*\(", header))
    names.discard("spde_plan")
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(_lib.lib, n), "libspde_b200.so does not export %s" % n
    assert _lib.lib.spde_abi_version() == 1
    # every declared function is bound with a signature in the Python layer as well
    assert names <= set(_lib.EXPORTS), names - set(_lib.EXPORTS)


def test_argument_errors_are_reported_without_a_gpu():
    from spdepy_b200 import _lib
    import pytest
    with pytest.raises(ValueError):
        _lib.PlanHandle(1, 5, 1, 3)            # mesh too small
    with pytest.raises(ValueError):
        _lib.PlanHandle(4, 4, 1, 2)            # periodic needs M,N >= 5


def test_no_oracle_import_in_product():
    """The product must never import the oracle (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "spdepy_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                for bad in ("spde_oracle", "plan_emulator", "cpu_cholesky", "stencil_oracle", "ref_harness", "import stencils"):
                    assert bad not in src, (f, bad)
