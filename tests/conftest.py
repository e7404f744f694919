import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture
def emu_device(monkeypatch):
    """Host-logic tests without a GPU: the ctypes device topology is replaced by the
    lane-emulation build of the same device core (tests only; see tests/helpers.py)."""
    from helpers import emu_solve
    from open_kinematics_b200 import _lib

    class EmuTopology:
        def __init__(self, program):
            self.program = program

        def close(self):
            pass

        def solve_batch(self, hardpoints, target_values, cfg=None, devices=None, want_positions=True,
                        want_tangents=False, want_metrics=False, want_design=False, params=None,
                        want_velocities=False, want_health=False, want_diagnostics=False, instance_targets=None,
                        want_worst_row=False, out=None, pinned=False):
            import numpy as np
            over = {} if cfg is None else {name: getattr(cfg, name) for name, _ in _lib.SolverCfg._fields_}
            tv = np.asarray(target_values, dtype=np.float64).reshape(len(self.program.target_points), -1)
            out = emu_solve(self.program, hardpoints, tv, params=params, want_health=want_health,
                            instance_targets=instance_targets, **over)
            if not want_metrics:
                out["metrics"] = None
            return out

    monkeypatch.setattr(_lib, "DeviceTopology", EmuTopology)
    monkeypatch.setattr(_lib, "require_device", lambda: None)
