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


def test_bank_conflict_model_on_known_patterns():
    import numpy as np
    """core/layout_tuning.py: wavefronts of 64-bit shared-memory accesses, half-warp at a time,
    16 bank pairs."""
    from open_kinematics_b200.core.layout_tuning import AccessTrace, UnitTrace, tune_block_slots, tune_unit_order

    def count(addresses):
        trace = AccessTrace(lb_base=0)
        trace.access(list(range(len(addresses))), [("ABS", a) for a in addresses])
        trace.freeze()
        return trace.wavefronts(np.arange(1)), trace.ideal()

    assert count(list(range(32))) == (2, 2)                 # consecutive doubles: one wavefront per half-warp
    assert count([7] * 32) == (2, 2)                        # broadcast
    assert count([16 * i for i in range(16)]) == (16, 1)    # same bank pair, 16 different doubles
    assert count([9 * i for i in range(16)]) == (1, 1)      # stride of a 3x3 block: conflict-free
    assert count([3 * i for i in range(16)]) == (1, 1)      # stride of a 3-vector
    assert count([8 * i for i in range(16)]) == (8, 1)      # stride 8: two bank pairs only
    # two blocks whose rows collide until one is moved
    trace = AccessTrace(lb_base=0)
    trace.access([0, 1], [("LB", 0), ("LB", 9 * 16)], (0, 1, 2))
    slot, before, after, ideal = tune_block_slots(trace, 17, iterations=200)
    assert (before, after, ideal) == (6, 3, 3) and sorted(slot) == list(range(17))
    units = UnitTrace(base=0, sizes=[3, 13, 3])
    units.access([0, 1], [(0, 0), (2, 0)], (0, 1, 2))      # units 0 and 2 start 16 doubles apart
    order, before, after, ideal = tune_unit_order(units, iterations=50)
    assert (before, after, ideal) == (6, 3, 3)
