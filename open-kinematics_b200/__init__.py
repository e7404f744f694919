"""B200-native batched suspension-kinematics solver (drop-in for the per-state
solve of nickmccleery/open-kinematics).

Host side: Python mirror of the reference's solve-path interface
(``core.solver.solve_suspension_sweep`` etc.).  Device side: hand-written
sm_100a CUDA behind the C ABI declared in ``include/okin.h`` and built from
``csrc/``.  There is no CPU fallback: every solve goes through the CUDA library.
"""

__version__ = "0.1.0"
