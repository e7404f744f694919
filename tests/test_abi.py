"""The C-ABI library loads on a CPU-only box and exports every symbol include/okin.h declares.
No compute call is made here."""

import ctypes
import os
import re

import pytest

from open_kinematics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "okin.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(okin_\w+)\s*\(", text)))


def test_header_declares_the_bound_functions():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_usage_errors_do_not_need_a_device():
    lib = _lib.load()
    cfg = _lib.default_cfg()
    assert cfg.step_tol == 1e-6 and cfg.coarse_tol == 1e-3 and cfg.fine_tol == 1e-4 and cfg.residual_tol == 1e-3 and cfg.max_iter == 50
    handle = ctypes.c_void_p()
    assert lib.okin_topology_create(None, ctypes.byref(handle)) == -1
    assert "null" in _lib.last_error()


def test_no_device_means_loud_failure():
    """The product path has no CPU fallback: without a device a solve raises."""
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()


def test_family_codes_agree_between_header_generator_and_compiler():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "generate_jacobians.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    from open_kinematics_b200.core.topology import FAMILY_CODE
    for code, fam in enumerate(gen.FAMILIES):
        assert FAMILY_CODE[fam.name] == code
    committed = open(os.path.join(ROOT, "open-kinematics_b200", "csrc", "okin_gen_constraints.cuh")).read()
    assert committed == gen.generate(), "csrc/okin_gen_constraints.cuh is stale: run tools/generate_jacobians.py"
