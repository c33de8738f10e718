"""The C-ABI library loads on a CPU-only box and exports every symbol include/altro_b200.h declares.
No compute call is made (there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import pytest

from altro_mpc_icra2021_b200 import solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "altro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(altro_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(solver.LIB_PATH):
        pytest.fail(f"{solver.LIB_PATH} missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(solver.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(solver.ABI_SYMBOLS) == syms, "python binding list and header disagree"


def test_default_options_match_reference_defaults():
    lib = solver.load_library()
    o = solver.AltroOpts()
    assert lib.altro_default_options(ctypes.byref(o)) == 0
    from altro_mpc_icra2021_b200.problem import SolverOptions

    d = SolverOptions()
    for name, _ in solver.AltroOpts._fields_:
        assert getattr(o, name) == getattr(d, name), name


def test_create_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = solver.load_library()
    h = ctypes.c_void_p()
    rc = lib.altro_create(ctypes.byref(h), 0, 6, 3, 21, 4, 0.05)
    assert rc != 0 and b"no CPU fallback" in lib.altro_last_error(None)
