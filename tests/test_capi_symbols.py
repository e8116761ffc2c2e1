"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol the header declares."""
import ctypes as C
import re

from optimistix_b200 import _lib, models
from optimistix_b200._solution import RESULTS


def _declared():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200opt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/b200opt.h but not exported"
    assert set(names) == set(_lib._SIGNATURES), set(names) ^ set(_lib._SIGNATURES)


def test_version_and_result_names(lib):
    assert lib.b200opt_version() == 100
    for code in RESULTS:
        assert lib.b200opt_result_name(code).decode() == RESULTS.name(code)
    # oracle, library and Python surface agree on names AND codes
    from oracle.solvers import RESULT_NAMES

    assert tuple(RESULT_NAMES) == RESULTS.names


def test_model_registry_matches_library(lib):
    models.check_against_library(lib)
    assert lib.b200opt_model_id(b"no_such_model") == -1


def test_default_options(lib):
    o = _lib.Options()
    lib.b200opt_default_options(C.byref(o))
    assert (o.rtol, o.atol, o.norm, o.max_steps, o.cauchy_termination, o.kappa) == (1e-8, 1e-8, 0, 256, 1, 1e-2)


def test_bad_arguments_fail_loudly(lib):
    o = _lib.make_options(1e-8, 1e-8)
    code = lib.b200opt_least_squares(0, 99, 1, 4, 3, 64, 1, 1, 1, 64, C.byref(o), 1, 1, 1, None, None, None)
    assert code == -1 and b"model" in lib.b200opt_last_error()
    code = lib.b200opt_least_squares(4, 0, 1, 4, 3, 64, 1, 1, 1, 64, C.byref(o), 1, 1, 1, None, None, None)
    assert code == -1 and b"solver" in lib.b200opt_last_error()
    code = lib.b200opt_least_squares(0, 0, 7, 4, 3, 64, 1, 1, 1, 64, C.byref(o), 1, 1, 1, None, None, None)
    assert code == -1 and b"dtype" in lib.b200opt_last_error()
    # batch == 0 is a valid no-op (empty input) and touches no pointer
    assert lib.b200opt_least_squares(0, 0, 1, 0, 3, 64, None, None, None, 64, C.byref(o), None, None, None, None, None, None) == 0


def test_xla_shim_uses_only_declared_entry_points_with_matching_arity():
    """integration/b200opt_xla.cc (the jax.ffi binding a maintainer would build; not compilable here) must call only functions the
    header declares, with the declared number of arguments."""
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = open(os.path.join(root, "integration", "b200opt_xla.cc")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)

    def arity(text, name):
        i = text.index(name + "(") + len(name) + 1
        depth, n, seen = 1, 0, False
        while depth:
            c = text[i]
            depth += c == "("
            depth -= c == ")"
            if depth == 1 and c == ",":
                n += 1
            seen = seen or not c.isspace()
            i += 1
        return n + 1 if seen and text[text.index(name + "(") + len(name) + 1 : i - 1].strip() not in ("", "void") else 0

    used = sorted(set(re.findall(r"\b(b200opt_[a-z0-9_]+)\s*\(", shim)))
    assert {"b200opt_least_squares", "b200opt_lsq_step", "b200opt_root_step", "b200opt_bfgs_step", "b200opt_lu_solve"} <= set(used)
    for name in used:
        assert name in _declared(), name
        assert arity(shim, name) == arity(hdr, name), (name, arity(shim, name), arity(hdr, name))


def test_run_time_family_needs_a_device(lib):
    """No CPU fallback: registering a run-time family without a CUDA device fails loudly instead of deferring."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a device is present")
    code = lib.b200opt_register_curve_family(b"nodev", 2, 1, b"struct Pre {};", _lib.CSRC_DIR.encode())
    assert code < 0 and lib.b200opt_last_error()
