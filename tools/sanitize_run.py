# small batch through every output path, for compute-sanitizer
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from helpers import *
from test_emu_core import _nominal
from open_kinematics_b200 import _lib
from open_kinematics_b200.core.topology import compile_suspension
for case in ("c3_rocker_ubar_coilover_roll", "c4_tbar_heave_shim_bump", "c2_macpherson_bump_steer"):
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    prog = compile_suspension(sus, sweep)
    rng=np.random.default_rng(0)
    hp = np.repeat(_nominal(sus, prog), 40, axis=0) + rng.normal(0,0.2,size=(40,3*prog.n_in))
    if case.startswith("c2"): hp = np.repeat(_nominal(sus, prog), 40, axis=0)
    topo = _lib.DeviceTopology(prog)
    v = arr["sweep_values"][:, :8]
    lean = topo.solve_batch(hp, v)
    full = topo.solve_batch(hp, v, want_metrics=True, want_velocities=True, want_health=True, want_diagnostics=True, want_tangents=True, want_design=True)
    print(case, (lean["status"]==0).mean(), (full["status"]==0).mean(), flush=True)
    topo.close()
