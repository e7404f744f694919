#!/usr/bin/env python3
"""Device-resident launches of the sweep kernel on the flagship workload, for ncu captures and A/B runs.

    python tools/kernel_probe.py launch [--instances N] [--reps R] [--full] [--sigma MM]
        R launches of N instances (what `ncu -k regex:okin_sweep_kernel -s 1 -c 1` captures;
        the first lean launch of a topology also runs the family calibration unless OKIN_LEAN_REGS is set)
    python tools/kernel_probe.py rates [--instances N]
        M states/s of the lean kernel and of the full kernel with all metric columns
    python tools/kernel_probe.py sigma [--instances N]
        lean rate on batches perturbed with sigma = 0.5 / 2 / 5 / 10 mm (the last has failing instances)

Environment: OKIN_LIB=<alternative build of libokin.so>, OKIN_LEAN_REGS=128|168, OKIN_WARPS_PER_CTA=<w>,
OKIN_TUNE=0 for the untuned shared-memory layout.  The first process on a fresh box measures 5-10 % low:
start every A/B session with a throw-away run."""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from open_kinematics_b200 import _lib  # noqa: E402
from open_kinematics_b200.core.sweep import BatchSolver  # noqa: E402


class Probe:
    def __init__(self, instances: int):
        sus, sweep = bench.workload_case()
        self.solver = BatchSolver(sus, sweep, tune_layout=bool(int(os.environ.get("OKIN_TUNE", "1"))))
        self.prog = self.solver.program
        self.n = instances
        self.nominal = self.solver.nominal_hardpoints()
        self.tv = torch.tensor(self.solver.values, device="cuda")
        self.steps = self.tv.shape[1]
        n, s, prog = instances, self.steps, self.prog
        self.pos = torch.empty((n, s, 3 * prog.n_out), device="cuda", dtype=torch.float64)
        self.status = torch.empty(n, device="cuda", dtype=torch.int32)
        self.failed = torch.empty_like(self.status)
        self.iters = torch.empty((n, s), device="cuda", dtype=torch.int32)
        self.maxres = torch.empty((n, s), device="cuda", dtype=torch.float64)
        self.metrics = None
        self.lib, self.cfg = _lib.load(), _lib.default_cfg()
        self.set_sigma(bench.SIGMA_MM)

    def set_sigma(self, sigma: float):
        h = bench.make_hardpoints_numpy(self.nominal, self.prog, self.n, 2)
        h = self.nominal[None, :] + (h - self.nominal[None, :]) * (sigma / bench.SIGMA_MM)
        self.hp = torch.tensor(h, device="cuda")

    def io(self, full: bool):
        if full and self.metrics is None:
            self.metrics = torch.empty((self.n, self.steps, len(self.prog.metric_names)), device="cuda", dtype=torch.float64)
        return _lib.BatchIO.of(hardpoints=self.hp.data_ptr(), target_values=self.tv.data_ptr(), positions=self.pos.data_ptr(),
                               status=self.status.data_ptr(), failed_step=self.failed.data_ptr(), iters=self.iters.data_ptr(),
                               max_residual=self.maxres.data_ptr(), metrics=self.metrics.data_ptr() if full else None)

    def launch(self, io):
        _lib.check(self.lib.okin_solve_batch_device(self.solver.topology.handle, ctypes.byref(self.cfg), 0, None,
                                                    self.n, self.steps, ctypes.byref(io)), "okin_solve_batch_device")

    def rate(self, full: bool, reps: int = 3) -> float:
        io = self.io(full)
        self.launch(io)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            self.launch(io)
        e1.record()
        torch.cuda.synchronize()
        return self.n * self.steps / (e0.elapsed_time(e1) / reps) / 1e3

    def summary(self) -> str:
        ok = self.status == 0
        return (f"ok fraction {float(ok.double().mean()):.4f} status counts {torch.bincount(self.status, minlength=4).tolist()} "
                f"mean nfev {float(self.iters.double().mean()):.3f} max nfev {int(self.iters.max())}")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("mode", choices=("launch", "rates", "sigma"))
    ap.add_argument("--instances", type=int, default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--sigma", type=float, default=bench.SIGMA_MM)
    a = ap.parse_args()
    lib = os.environ.get("OKIN_LIB", "in-tree")
    if a.mode == "launch":
        p = Probe(a.instances or 32768)
        p.set_sigma(a.sigma)
        io = p.io(a.full)
        for _ in range(a.reps):
            p.launch(io)
        torch.cuda.synchronize()
        print("full" if a.full else "lean", "sigma", a.sigma, p.summary(), p.solver.topology.launch_geometry(p.n), flush=True)
    elif a.mode == "rates":
        p = Probe(a.instances or (1 << 18))
        for label, full in (("lean", False), ("metrics", True)):
            print(lib, label, "M states/s %.2f" % p.rate(full), p.summary(), flush=True)
    else:
        p = Probe(a.instances or (1 << 17))
        for sigma in (0.5, 2.0, 5.0, 10.0):
            p.set_sigma(sigma)
            print(lib, "sigma", sigma, "M sweep states/s (all) %.2f" % p.rate(False, reps=2), p.summary(), flush=True)


if __name__ == "__main__":
    main()
