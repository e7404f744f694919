#!/usr/bin/env python3
"""BASELINE.json configs[3] and configs[4] at their stated sizes, on 1..8 GPUs.

    python tools/config_scale.py --config c4 [--instances 10000000] [--steps 101]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/config_scale.py --config c5

c4: double-wishbone axle with torsion bars, T-bar ARB, rocker-to-rocker heave link and camber shims,
    101-step bump and 101-step roll sweeps, 1e7 Monte-Carlo tolerance instances (hardpoints sigma
    0.25 mm, shim set-up thickness drawn per instance) in total over the ranks.
c5: DoE sensitivity sweep over a hardpoint grid: 1e8 instance-steps (4 761 905 instances x 21 steps of
    the flagship axle, each instance one grid node of a 6-factor design) with every metric column
    (motion ratios, roll-centre metrics, ...) evaluated on the device.

The instance range of a rank is processed in chunks whose outputs stay on the device (1e7 x 101
states of all-point positions would be 1.1 TB): per chunk the positions / metrics are reduced on the
device to what a tolerance study keeps (status counts, min / max / mean of selected metric columns),
and the chunk buffers are reused.  One JSON line per run (rank 0): whole-job states/s, timed with
CUDA events on the launch stream, max over ranks.  No data-path collective: ranks only reduce the
timing and the summary statistics at the end.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import build_case, load_golden  # noqa: E402
from open_kinematics_b200 import _lib  # noqa: E402
from open_kinematics_b200.core.input import build_sweep  # noqa: E402
from open_kinematics_b200.core.sweep import BatchSolver  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))
from config_throughput import axle_sweep  # noqa: E402


def doe_levels(rank_begin: int, count: int, n_factors: int, levels: int, device) -> torch.Tensor:
    """Grid coordinates in [-1, 1] of DoE nodes [rank_begin, rank_begin + count): node index written
    in base ``levels`` (a full factorial design walked in order, wrapped when exhausted)."""
    idx = torch.arange(rank_begin, rank_begin + count, device=device, dtype=torch.int64)
    cols = []
    for _ in range(n_factors):
        cols.append((idx % levels).double() / (levels - 1) * 2.0 - 1.0)
        idx = idx // levels
    return torch.stack(cols, dim=1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["c4", "c5"], required=True)
    ap.add_argument("--instances", type=int, default=None, help="total over all ranks")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--mode", default="bump", choices=["bump", "roll"], help="c4 sweep kind")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()

    if args.config == "c4":
        meta, _ = load_golden("c4_tbar_heave_shim_bump")
        sus, _ = build_case(meta)
        steps = args.steps or 101
        sweep = build_sweep(axle_sweep(steps, 50.0, args.mode == "roll"), sus)
        total = args.instances or 10_000_000
        want_metrics, sigma = False, 0.25
        label = f"c4_tbar_heave_shim_{args.mode}{steps}_montecarlo"
    else:
        meta, _ = load_golden("c3_rocker_ubar_coilover_roll")
        sus, sweep = build_case(meta)
        steps = sweep.n_steps
        total = args.instances or -(-100_000_000 // steps)      # 1e8 instance-steps
        want_metrics, sigma = True, 0.0
        label = "c5_doe_grid_c3_axle_roll21_all_metrics"
    solver = BatchSolver(sus, sweep, tune_layout=True)
    prog, topo = solver.program, solver.topology
    begin, count = _lib.shard_range(total, rank, world)
    nominal = torch.tensor(solver.nominal_hardpoints(), device=dev, dtype=torch.float64)
    own, pairs = bench.perturbation_mask(prog)
    own_t = torch.tensor(own, device=dev)
    p_left, p_right = torch.tensor(pairs[:, 0], device=dev), torch.tensor(pairs[:, 1], device=dev)
    flip = torch.tensor([1.0, -1.0, 1.0], device=dev, dtype=torch.float64)
    # centre-line points keep y = 0 (T-bar pivot: the reference's build validation, axle/mechanisms.py:626-643)
    from open_kinematics_b200.core.primitives.point_ref import Side
    centre = [i for i, k in enumerate(prog.in_keys) if getattr(k, "side", None) is Side.CENTER
              and abs(float(solver.nominal_hardpoints()[3 * i + 1])) < 1e-9]
    centre_t = torch.tensor(centre, device=dev, dtype=torch.int64)
    S, nin, nout, nm = steps, prog.n_in, prog.n_out, len(prog.metric_names)
    chunk = min(args.chunk, max(count, 1))
    # chunk buffers, reused
    hp = torch.empty((chunk, nin, 3), device=dev, dtype=torch.float64)
    par = None
    if prog.param_names:
        par = torch.tensor(prog.param_default, device=dev, dtype=torch.float64).repeat(chunk, 1)
        shim_cols = torch.tensor([i for i, nme in enumerate(prog.param_names) if nme.endswith("setup_thickness")], device=dev)
    tv = torch.tensor(solver.values, device=dev, dtype=torch.float64).contiguous()
    pos = torch.empty((chunk, S, 3 * nout), device=dev, dtype=torch.float64) if not want_metrics else None
    met = torch.empty((chunk, S, nm), device=dev, dtype=torch.float64) if want_metrics else None
    status = torch.empty(chunk, device=dev, dtype=torch.int32)
    failed = torch.empty(chunk, device=dev, dtype=torch.int32)
    iters = torch.empty((chunk, S), device=dev, dtype=torch.int32)
    maxres = torch.empty((chunk, S), device=dev, dtype=torch.float64)
    io = _lib.BatchIO.of(hardpoints=hp.data_ptr(), params=par.data_ptr() if par is not None else None,
                         target_values=tv.data_ptr(), positions=pos.data_ptr() if pos is not None else None,
                         metrics=met.data_ptr() if met is not None else None, status=status.data_ptr(),
                         failed_step=failed.data_ptr(), iters=iters.data_ptr(), max_residual=maxres.data_ptr())
    lib, cfg = _lib.load(), _lib.default_cfg()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # DoE factors: six left-side hardpoint coordinates moved +-2 mm (c5); Monte Carlo: all of them (c4)
    n_factors, levels = 6, 13                                  # 13^6 = 4.83e6 grid nodes >= 4.76e6 instances
    factor_slots = [(int(own[k % own.size]), k % 3) for k in range(0, 6 * 5, 5)]

    def fill(lo: int, c: int) -> None:
        hp[:c] = nominal.reshape(1, nin, 3)
        if args.config == "c4":
            hp[:c, own_t, :] += sigma * torch.randn((c, own.size, 3), device=dev, dtype=torch.float64, generator=gen)
            par[:c, shim_cols] = 29.5 + torch.rand((c, shim_cols.numel()), device=dev, dtype=torch.float64, generator=gen)
        else:
            grid = doe_levels(begin + lo, c, n_factors, levels, dev)
            for f, (slot, axis) in enumerate(factor_slots):
                hp[:c, slot, axis] += 2.0 * grid[:, f]
        if centre:
            hp[:c, centre_t, 1] = 0.0
        hp[:c, p_right, :] = hp[:c, p_left, :] * flip

    def launch(c: int) -> None:
        _lib.check(lib.okin_solve_batch_device(topo.handle, ctypes.byref(cfg), local, ctypes.c_void_p(
            torch.cuda.current_stream().cuda_stream), c, S, ctypes.byref(io)), label)

    n_ok = torch.zeros((), device=dev, dtype=torch.int64)
    accepted = torch.zeros((), device=dev, dtype=torch.int64)
    nfev = torch.zeros((), device=dev, dtype=torch.int64)
    stat_cols = [prog.metric_names.index(k) for k in prog.metric_names
                 if k in ("roll_center_z", "camber_left", "deriv_damper_length_wrt_hub_z_left",
                          "deriv_arb_twist_wrt_hub_z_left")] if want_metrics else []
    col_min = torch.full((len(stat_cols),), float("inf"), device=dev, dtype=torch.float64)
    col_max = torch.full((len(stat_cols),), float("-inf"), device=dev, dtype=torch.float64)
    col_idx = torch.tensor(stat_cols, device=dev, dtype=torch.int64)

    def process(lo: int, c: int, events) -> None:
        """One chunk: inputs on the device, solve, reduce what a tolerance study keeps."""
        nonlocal n_ok, accepted, nfev, col_min, col_max
        fill(lo, c)
        events[0].record()
        launch(c)
        events[1].record()
        ok = status[:c] == 0
        n_ok += ok.sum()
        accepted += torch.where(ok, torch.full_like(failed[:c], S), failed[:c].clamp(min=0)).sum()
        nfev += iters[:c].sum()
        if stat_cols:
            sel = met[:c].index_select(2, col_idx)                                   # [c, S, k]
            good = ok[:, None, None] & ~torch.isnan(sel)
            col_min = torch.minimum(col_min, torch.where(good, sel, float("inf")).amin(dim=(0, 1)))
            col_max = torch.maximum(col_max, torch.where(good, sel, float("-inf")).amax(dim=(0, 1)))

    # warm-up on one chunk (untimed: kernel-family calibration, torch kernel loading), counters reset after
    warm = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(2):
        process(0, chunk, warm)
    torch.cuda.synchronize()
    n_ok.zero_(), accepted.zero_(), nfev.zero_()
    col_min.fill_(float("inf")), col_max.fill_(float("-inf"))
    if world > 1:
        dist.barrier()
    n_chunks = (count + chunk - 1) // chunk
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n_chunks)]
    wall0, wall1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0.record()
    for k, lo in enumerate(range(0, count, chunk)):
        process(lo, min(chunk, count - lo), ev[k])           # no host synchronisation inside the loop
    wall1.record()
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = n_chunks
    wall_ms = wall0.elapsed_time(wall1)
    t = torch.tensor([kernel_ms, wall_ms], device=dev, dtype=torch.float64)
    sums = torch.stack([n_ok, accepted, nfev]).double()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        if stat_cols:
            dist.all_reduce(col_min, op=dist.ReduceOp.MIN)
            dist.all_reduce(col_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        k_ms, w_ms = float(t[0]), float(t[1])
        n_ok_all, accepted_all, nfev_all = (float(v) for v in sums)
        line = {
            "config": label, "n_gpus": world, "instances_total": total, "sweep_steps": S,
            "instance_steps_total": total * S, "chunk_instances": chunk, "metrics_on_device": bool(want_metrics),
            "n_metric_columns": nm if want_metrics else 0, "n_unknowns": prog.n_unknowns,
            "states_per_s": accepted_all / (k_ms * 1e-3),
            "states_per_s_incl_input_generation_and_reductions": accepted_all / (w_ms * 1e-3),
            "kernel_seconds_max_over_ranks": k_ms * 1e-3, "wall_seconds_max_over_ranks": w_ms * 1e-3,
            "ok_fraction": n_ok_all / total, "accepted_states": accepted_all,
            "mean_nfev_per_state": nfev_all / max(accepted_all, 1.0), "launches_per_rank": launches,
            "outputs": "metrics (all columns) + nfev + max_residual + status, reduced on the device per chunk"
                       if want_metrics else "positions (all points) + nfev + max_residual + status, kept on the device per chunk",
            "launch": topo.launch_geometry(chunk, local), "scaling": "strong (total work fixed)",
        }
        if stat_cols:
            line["metric_ranges"] = {prog.metric_names[c]: [float(col_min[i]), float(col_max[i])]
                                     for i, c in enumerate(stat_cols)}
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
