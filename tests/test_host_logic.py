"""Host-side logic that needs no GPU: the options struct against the C library's defaults, the linear_solver / data_dtype
mapping, the solver protocol surface, the bench's reference arm contract."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import optimistix_b200 as optx
from optimistix_b200 import _lib
from optimistix_b200._entry import _data16_kind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_options_struct_matches_the_library():
    """`_lib.Options` must be laid out like `b200opt_options_t`: the library writes its defaults into a poisoned struct."""
    lib = _lib.load()
    raw = (C.c_ubyte * (C.sizeof(_lib.Options) + 16))(*([0xAB] * (C.sizeof(_lib.Options) + 16)))
    o = _lib.Options.from_buffer(raw)
    lib.b200opt_default_options(C.byref(o))
    assert (o.rtol, o.atol, o.norm, o.max_steps, o.cauchy_termination, o.root_near_zero, o.kappa) == (1e-8, 1e-8, 0, 256, 1, 0, 1e-2)
    assert (o.linear_solver, o.data_dtype) == (_lib.LINSOLVE_NORMAL_CHOLESKY, _lib.DATA_AS_COMPUTE)
    assert o.scratch is None  # NULL: the library takes the queue counter from the stream-ordered allocator
    assert all(b == 0xAB for b in bytes(raw)[C.sizeof(_lib.Options):]), "the library wrote past the struct the binding declares"
    with open(_lib.HEADER_PATH) as fh:
        header = fh.read()
    for name, value in (("B200OPT_LINSOLVE_QR", _lib.LINSOLVE_QR), ("B200OPT_DATA_F16", _lib.DATA_F16), ("B200OPT_DATA_U16", _lib.DATA_U16),
                        ("B200OPT_DATA_I16", _lib.DATA_I16)):
        assert f"#define {name} {value}" in header


def test_linear_solver_tokens():
    assert _lib.linear_solver_id(optx.NormalCholesky()) == _lib.LINSOLVE_NORMAL_CHOLESKY
    assert _lib.linear_solver_id(optx.Cholesky()) == _lib.LINSOLVE_NORMAL_CHOLESKY
    assert _lib.linear_solver_id(optx.AutoLinearSolver(well_posed=None)) == _lib.LINSOLVE_NORMAL_CHOLESKY
    assert _lib.linear_solver_id(optx.QR()) == _lib.LINSOLVE_QR
    assert _lib.linear_solver_id(None) == _lib.LINSOLVE_NORMAL_CHOLESKY
    for bad in (optx.LU(), optx.AutoLinearSolver(well_posed=True), object()):
        with pytest.raises(NotImplementedError, match="linear_solver"):
            _lib.linear_solver_id(bad)
    assert optx.LevenbergMarquardt(1e-3, 1e-3).descent.linear_solver == optx.NormalCholesky()
    assert optx.LevenbergMarquardt(1e-3, 1e-3, linear_solver=optx.QR()).descent.linear_solver == optx.QR()
    o = _lib.make_options(1e-3, 1e-3, "max", 7, linear_solver=optx.QR(), data_dtype=_lib.DATA_U16)
    assert (o.linear_solver, o.data_dtype, o.max_steps) == (1, 2, 7)


def test_sixteen_bit_data_detection():
    import torch

    assert _data16_kind(np.zeros(3, np.float16)) == _lib.DATA_F16
    assert _data16_kind(np.zeros(3, np.uint16)) == _lib.DATA_U16
    assert _data16_kind(np.zeros(3, np.int16)) == _lib.DATA_I16
    assert _data16_kind(np.zeros(3, np.float32)) == _lib.DATA_AS_COMPUTE
    assert _data16_kind(torch.zeros(3, dtype=torch.float16)) == _lib.DATA_F16
    assert _data16_kind(torch.zeros(3, dtype=torch.int16)) == _lib.DATA_I16
    assert _data16_kind(torch.zeros(3, dtype=torch.float64)) == _lib.DATA_AS_COMPUTE
    assert _data16_kind([1.0, 2.0]) == _lib.DATA_AS_COMPUTE


def test_solver_classes_carry_the_reference_protocol():
    """init / step / terminate / postprocess with the reference's parameter names (optimistix/_iterate.py:33-184)."""
    import inspect

    want = {"init": ["fn", "y", "args", "options", "f_struct", "aux_struct", "tags"],
            "step": ["fn", "y", "args", "options", "state", "tags"],
            "terminate": ["fn", "y", "args", "options", "state", "tags"],
            "postprocess": ["fn", "y", "aux", "args", "options", "state", "tags", "result"]}
    for cls in (optx.LevenbergMarquardt, optx.GaussNewton, optx.Newton, optx.Chord, optx.BFGS, optx.DFP):
        solver = cls(1e-3, 1e-3)
        for name, params in want.items():
            assert list(inspect.signature(getattr(solver, name)).parameters) == params, (cls.__name__, name)
        assert solver.postprocess(None, 1, 2, None, {}, None, frozenset(), None) == (1, 2, {})
    # verbose: False -> silent (None), True -> the default printer, a callable -> itself (gauss_newton.py:382-395)
    assert optx.LevenbergMarquardt(1e-3, 1e-3).verbose is None
    assert callable(optx.BFGS(1e-3, 1e-3, verbose=True).verbose)
    f = lambda **kw: None  # noqa: E731
    assert optx.GaussNewton(1e-3, 1e-3, verbose=f).verbose is f
    with pytest.raises(TypeError):
        optx.DFP(1e-3, 1e-3, verbose="yes")


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU and prints the contract line: impl, metric, value, a cpu_baseline that
    describes this run, e2e with zero copy bytes, and the number of problems one timed step of the arm solves."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "exp_decay", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "problems/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "problems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["problems_per_timed_step_of_this_arm"] == 4096
    # ranks other than 0 of a torchrun launch print nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
