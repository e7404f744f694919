#!/usr/bin/env python3
"""Headline benchmark: Newton-solved sweep states per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--instances I] [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric's target is quoted on):
mirrored double-wishbone axle with pushrod-rocker coilovers and U-bar ARB (reference
tests/data/axle_geometry_rocker.yaml + coilover points, SURVEY.md section 8d "C3"), 21-step
roll sweep +-20 mm with the rack held, I hardpoint-perturbed instances per GPU (sigma 0.5 mm
on the left + centre points, right side mirrored; seeded).  One bench "step" = one pass of the
hot path over the whole batch = I x 21 solved states per GPU.

Own arm   : CUDA path.  `value` = states/s with inputs resident in HBM (device entry point of the
            C ABI, CUDA events on the launch stream); `e2e` = same metric through the host-buffer
            C-ABI call (pinned host buffers, H2D + kernel + D2H inside the timed region).
Reference : `--impl reference` times the oracle port of the reference's CPU algorithm
            (oracle/solve.py: SciPy MINPACK LM on the reference's residual/Jacobian callbacks)
            on all host cores over a bounded sample of the same workload.
Multi-GPU : torchrun, one rank per GPU, instance ranges sharded with no data-path collective
            (weak scaling); timing = max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "solved_sweep_states_per_sec"
UNIT = "states/s"
WORKLOAD = "c3_rocker_ubar_coilover_axle_roll21"
N_STEPS_SWEEP = 21
SIGMA_MM = 0.5


def workload_case():
    """Geometry + sweep of the flagship workload from the committed golden inputs (the GPU box
    has no /root/reference)."""
    from helpers import build_case, load_golden
    meta, _ = load_golden("c3_rocker_ubar_coilover_roll")
    return build_case(meta)


def perturbation_mask(program) -> tuple:
    """(index arrays) so that left + centre points get independent noise and every right point
    mirrors its left twin (y -> -y), as the reference mirrors omitted right sides (build.py:344-354)."""
    from open_kinematics_b200.core.primitives.point_ref import PointRef, Side
    slot = {k: i for i, k in enumerate(program.in_keys)}
    own = [i for k, i in slot.items() if k.side in (Side.LEFT, Side.CENTER)]
    pairs = [(slot[PointRef(Side.LEFT, k.point)], i) for k, i in slot.items() if k.side is Side.RIGHT]
    return np.array(own), np.array(pairs)


def make_hardpoints_numpy(nominal: np.ndarray, program, n: int, seed: int) -> np.ndarray:
    own, pairs = perturbation_mask(program)
    rng = np.random.default_rng(seed)
    hp = np.repeat(nominal[None, :].reshape(1, -1, 3), n, axis=0)
    hp[:, own, :] += rng.normal(0.0, SIGMA_MM, size=(n, own.size, 3))
    hp[:, pairs[:, 1], :] = hp[:, pairs[:, 0], :] * np.array([1.0, -1.0, 1.0])
    return hp.reshape(n, -1)


def algorithmic_flops_per_state(stats: dict, mean_iters: float, n_targets: int) -> float:
    """Executed-sparse flop model: per linear solve 2*(assembly + update
    + triangular solves) FMAs + ~row evaluation; plus the tangent solves per state."""
    eval_flops = 60.0 * stats["n_rows"]                     # ~45 flops per distance row + a few heavy rows
    per_iter = 2.0 * (stats["asm_fma"] + stats["g_fma"] + stats["update_fma"] + stats["solve_fma"]) \
        + 11.0 * stats["scale_tasks"] + 40.0 * stats["n_free"] + eval_flops   # row scalings + 3x3 Choleskys
    lin_solves = max(mean_iters - 1.0, 1.0)                 # nfev counts one residual-only evaluation
    return lin_solves * per_iter + eval_flops + n_targets * 2.0 * stats["solve_fma"]


def bind_to_gpu_numa_node(index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to GPU ``index`` so that the pinned host
    buffers of the end-to-end path are first-touched on the GPU's own NUMA node (with 8 ranks the
    D2H stream otherwise crosses the socket interconnect).  Returns the CPU count, 0 if unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 - affinity is an optimisation only
        return 0


def rank_seed(rank: int) -> int:
    """Perturbation seed of a rank: config index (2) + rank, so ranks draw disjoint streams."""
    return 2 + rank


def reduce_max_ms(values, device, world: int) -> list:
    """Max over ranks of per-rank timings (the job is as slow as its slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def job_throughput(units_per_rank: float, world: int, ms: float) -> float:
    """Whole-job units per second with per-rank work fixed (weak scaling)."""
    return world * units_per_rank / (ms * 1e-3)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def oracle_instance_worker(args):
    """Solve one perturbed instance with the oracle port (runs in a worker process)."""
    seed, = args
    from helpers import authored_positions, oracle_problem
    sus, sweep = workload_case()
    problem, values = oracle_problem(sus, sweep)
    from open_kinematics_b200.core.primitives.point_ref import Side
    from oracle.solve import solve_sweep
    rng = np.random.default_rng(seed)
    auth = authored_positions(sus)
    pert = {}
    for k, v in auth.items():
        if k.side in (Side.LEFT, Side.CENTER):
            pert[k] = v + rng.normal(0.0, SIGMA_MM, 3)
    for k, v in auth.items():
        if k.side is Side.RIGHT:
            twin = type(k)(Side.LEFT, k.point)
            pert[k] = pert[twin] * np.array([1.0, -1.0, 1.0]) if twin in pert else v
    out = solve_sweep(problem, pert, values)
    return int(out["status"] == 0) * values.shape[1]


def cpu_baseline(sample_instances: int, processes: int) -> dict:
    """Oracle port timed on the host cores over a bounded sample of the workload."""
    seeds = [(1000 + i,) for i in range(sample_instances)]
    t0 = time.perf_counter()
    if processes <= 1:
        solved = sum(oracle_instance_worker(s) for s in seeds)
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(processes) as pool:
            solved = sum(pool.map(oracle_instance_worker, seeds, chunksize=1))
    dt = time.perf_counter() - t0
    return {"value": solved / dt, "unit": UNIT, "cores": processes, "kind": "port",
            "sample": f"{sample_instances} perturbed instances x {N_STEPS_SWEEP} steps of {WORKLOAD}, "
                      f"oracle/solve.py (SciPy MINPACK LM, reference default tolerances), {dt:.1f} s"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 4)
    # warm-up (imports, process pool start) then K timed steps of `per_step` instances each
    times, solved = [], 0
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            seeds = [(5000 + step * per_step + i,) for i in range(per_step)]
            t0 = time.perf_counter()
            got = sum(pool.map(oracle_instance_worker, seeds, chunksize=1))
            dt = time.perf_counter() - t0
            if step >= min(args.warmup, 1):   # one warm-up step is enough for a CPU pool
                times.append(dt)
                solved += got
    value = solved / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_step": per_step, "sweep_steps": N_STEPS_SWEEP,
                   "sigma_mm": SIGMA_MM},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} instances x {N_STEPS_SWEEP} steps per bench step, "
                                   f"multiprocessing.Pool({cores}), oracle/solve.py (SciPy MINPACK LM)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_cuda(args) -> None:
    import ctypes

    import torch
    import torch.distributed as dist

    from open_kinematics_b200 import _lib
    from open_kinematics_b200.core.sweep import BatchSolver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_device()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0

    sus, sweep = workload_case()
    solver = BatchSolver(sus, sweep, tune_layout=True)     # bank-conflict-aware block placement (compile time)
    prog, topo = solver.program, solver.topology
    n_inst, S = args.instances, N_STEPS_SWEEP
    nin3, nout3, nt, n = 3 * prog.n_in, 3 * prog.n_out, len(prog.target_points), prog.n_unknowns
    lib = _lib.load()
    cfg = _lib.default_cfg()

    # ---- device-resident inputs (value) ---------------------------------------------------
    own, pairs = perturbation_mask(prog)
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank_seed(rank))
    nominal = torch.tensor(solver.nominal_hardpoints(), device=dev, dtype=torch.float64).reshape(-1, 3)
    hp = nominal.unsqueeze(0).repeat(n_inst, 1, 1)
    own_t = torch.tensor(own, device=dev)
    hp[:, own_t, :] += SIGMA_MM * torch.randn((n_inst, own.size, 3), device=dev, dtype=torch.float64, generator=gen)
    flip = torch.tensor([1.0, -1.0, 1.0], device=dev, dtype=torch.float64)
    hp[:, torch.tensor(pairs[:, 1], device=dev), :] = hp[:, torch.tensor(pairs[:, 0], device=dev), :] * flip
    hp = hp.reshape(n_inst, nin3).contiguous()
    tv = torch.tensor(solver.values, device=dev, dtype=torch.float64).contiguous()
    pos = torch.empty((n_inst, S, nout3), device=dev, dtype=torch.float64)
    status = torch.empty(n_inst, device=dev, dtype=torch.int32)
    failed = torch.empty(n_inst, device=dev, dtype=torch.int32)
    iters = torch.empty((n_inst, S), device=dev, dtype=torch.int32)
    maxres = torch.empty((n_inst, S), device=dev, dtype=torch.float64)

    d_io = _lib.BatchIO.of(hardpoints=hp.data_ptr(), target_values=tv.data_ptr(), positions=pos.data_ptr(),
                           status=status.data_ptr(), failed_step=failed.data_ptr(), iters=iters.data_ptr(),
                           max_residual=maxres.data_ptr())

    def launch():
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.okin_solve_batch_device(
            topo.handle, ctypes.byref(cfg), local, ctypes.c_void_p(stream), n_inst, S, ctypes.byref(d_io)),
            "okin_solve_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        launch()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        launch()
        ev[k + 1].record()
    barrier()
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop() if rank == 0 else None

    ok_frac = float((status == 0).double().mean().item())
    mean_iters = float(iters.double().mean().item())
    states_per_launch = n_inst * S

    # ---- end to end through the host-buffer C-ABI call ------------------------------------
    import psutil
    avail = psutil.virtual_memory().available
    bytes_per_inst_out = S * nout3 * 8 + S * 12 + 8
    e2e_inst = int(min(n_inst, max(4096, min(30e9, 0.25 * avail / max(world, 1)) // bytes_per_inst_out)))
    h_hp = torch.empty((e2e_inst, nin3), dtype=torch.float64, pin_memory=True)
    h_hp.copy_(hp[:e2e_inst].cpu())
    h_tv = torch.tensor(solver.values, dtype=torch.float64).contiguous()
    h_pos = torch.empty((e2e_inst, S, nout3), dtype=torch.float64, pin_memory=True)
    h_status = torch.empty(e2e_inst, dtype=torch.int32, pin_memory=True)
    h_failed = torch.empty(e2e_inst, dtype=torch.int32, pin_memory=True)
    h_iters = torch.empty((e2e_inst, S), dtype=torch.int32, pin_memory=True)
    h_maxres = torch.empty((e2e_inst, S), dtype=torch.float64, pin_memory=True)
    devs = np.array([local], dtype=np.int32)

    h_io = _lib.BatchIO.of(hardpoints=h_hp.data_ptr(), target_values=h_tv.data_ptr(), positions=h_pos.data_ptr(),
                           status=h_status.data_ptr(), failed_step=h_failed.data_ptr(), iters=h_iters.data_ptr(),
                           max_residual=h_maxres.data_ptr())

    def e2e_call():
        _lib.check(lib.okin_solve_batch(topo.handle, ctypes.byref(cfg), e2e_inst, S, ctypes.byref(h_io),
                                        devs.ctypes.data, 1), "okin_solve_batch")

    for _ in range(2):
        e2e_call()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()                                              # synchronises internally
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_ok = float((h_status == 0).double().mean().item())

    # ---- reduce over ranks (max time) ------------------------------------------------------
    total_ms_max, e2e_s_max = reduce_max_ms([total_ms, e2e_s], dev, world)

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = job_throughput(states_per_launch, world, ms_per_step)
        e2e_value = job_throughput(e2e_inst * S, world, e2e_s_max * 1e3)
        peak = ctypes.c_double(0.0)
        _lib.check(lib.okin_fp64_peak(local, ctypes.byref(peak)), "okin_fp64_peak")
        model_flops_state = algorithmic_flops_per_state(prog.stats, mean_iters, nt)
        k_ms = float(np.mean(kernel_ms))
        executed = None
        try:
            executed = json.load(open(os.path.join(ROOT, "profiles", "fp64_flops.json")))["executed_fp64_flops_per_state"]
        except (OSError, KeyError):
            pass
        # a kernel that exploits the sparsity reports executed flops (ncu
        # 2*dfma + dmul + dadd per state, profiles/fp64_flops.json); the structural model is the fallback
        flops_state = executed if executed else model_flops_state
        achieved_tflops = flops_state * states_per_launch / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes_state = nout3 * 8 + 12 + (nin3 * 8 + 8) / S
        hbm_achieved = alg_bytes_state * states_per_launch / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            traffic = tj["dram_bytes_per_state"] * states_per_launch
            traffic_src = tj["source"]
        except (OSError, KeyError):
            pass
        smem_pipe = None
        try:
            smem_pipe = json.load(open(os.path.join(ROOT, "profiles", "smem_pipe.json")))
        except OSError:
            pass
        geo = topo.launch_geometry(n_inst, local)
        base = cpu_baseline(sample_instances=args.cpu_sample, processes=1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "instances_per_gpu": n_inst, "sweep_steps": S, "sigma_mm": SIGMA_MM,
                "n_unknowns": n, "n_rows": prog.stats["n_rows"], "ok_fraction": ok_frac,
                "mean_nfev_per_state": mean_iters, "outputs": "positions(all points)+nfev+max_residual+status",
                "l2_policy": f"inputs+outputs per launch {n_inst * (nin3 * 8 + S * nout3 * 8) / 1e9:.2f} GB >> 126 MB L2",
                "launch": geo, "layout_tuning": prog.stats.get("layout_tuning"), "numa_local_cpus": local_cpus, "e2e_instances_per_gpu": e2e_inst, "e2e_ok_fraction": e2e_ok,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_inst * nin3 * 8 + nt * S * 8),
                    "d2h_bytes_per_step": int(e2e_inst * bytes_per_inst_out)},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved_tflops, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved_tflops / peak.value if peak.value else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "okin_fp64_peak DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                "flops_per_state": flops_state, "flops_source": "ncu executed (2*dfma+dmul+dadd)" if executed
                else "structural model", "model_flops_per_state": model_flops_state,
                "dense_lu_equivalent_flops_per_state": float(prog.stats["dense_lu_flops"]) * max(mean_iters - 1.0, 1.0),
                "kernel_ms": k_ms,
                # the unit ncu shows closest to its peak is the shared-memory data pipe, not FP64
                "shared_memory_pipe_ncu": smem_pipe,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes_per_state": alg_bytes_state,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            },
            "cpu_baseline": base,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--instances", type=int, default=1 << 20, help="perturbed instances per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=8, help="instances in the CPU-baseline sample")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
