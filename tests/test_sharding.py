"""Multi-GPU plumbing that does not need a GPU: the instance-range rule of the C ABI and the
rank protocol of bench.py (world_size 2 over gloo on the CPU)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_kinematics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,g", [(0, 1), (1, 8), (7, 2), (1000003, 8), (1 << 20, 4), (5, 8)])
def test_shard_ranges_partition_the_batch(n, g):
    ranges = [_lib.shard_range(n, k, g) for k in range(g)]
    assert ranges[0][0] == 0
    assert sum(c for _, c in ranges) == n
    for (b0, c0), (b1, _) in zip(ranges, ranges[1:]):
        assert b0 + c0 == b1
    counts = [c for _, c in ranges]
    assert max(counts) - min(counts) <= 1


def test_shard_range_rejects_bad_arguments():
    with pytest.raises(RuntimeError):
        _lib.shard_range(10, 3, 2)


def _rank_main(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank owns a contiguous instance range and its own perturbation seed
    n_total = 1001
    begin, count = _lib.shard_range(n_total, rank, world)
    seed = bench.rank_seed(rank)
    # per-rank timing -> job timing = max over ranks; job throughput = all units / that time
    local_ms = 10.0 * (rank + 1)
    ms_max = bench.reduce_max_ms([local_ms, 2.0 * local_ms], torch.device("cpu"), world)
    value = bench.job_throughput(units_per_rank=count * 21, world=world, ms=ms_max[0])
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([begin, count, seed, ms_max[0], ms_max[1], value]))
    dist.barrier()
    dist.destroy_process_group()


def test_bench_rank_protocol_world_size_2_gloo(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    assert rows[0][0] == 0 and rows[0][0] + rows[0][1] == rows[1][0] and rows[1][0] + rows[1][1] == 1001
    assert rows[0][2] != rows[1][2]                      # different perturbation streams
    for r in rows:                                       # every rank sees the max over ranks
        assert r[3] == 20.0 and r[4] == 40.0


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ carry every key of the bench contract (so a change
    to bench.py that drops one shows up here, without a GPU)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("r01_bench_1gpu.json", "r01_bench_8gpu.json", "r02_bench_1gpu.json", "r02_bench_8gpu.json",
                 "r02_bench_8gpu_single_process.json"):
        line = json.loads(open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks",
                    "roofline", "cpu_baseline"):
            assert key in line, (name, key)
        assert line["dtype"] == "f64" and line["scaling"] == "weak" and line["vs_baseline"] is None
        assert line["higher_is_better"] is True and line["warmup"] >= 3 and "workload" in line["config"]
        assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
        assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
        assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
        accepted = line["config"].get("accepted_state_fraction", 1.0)      # round 2: only accepted states count
        assert abs(line["value"] - accepted * line["n_gpus"] * line["config"]["instances_per_gpu"]
                   * line["config"]["sweep_steps"] / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
        if name.startswith("r02"):
            assert line["config"]["parity_max_mm"] <= 1e-6 and line["config"]["parity_flags_identical"]
            assert set(line["e2e"]["variants"]) >= {"all_points", "metrics_only"}
            assert line["e2e"]["value"] == line["e2e"]["variants"]["all_points"]["value"]
    ref = json.loads(open(os.path.join(root, "profiles", "r01_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "port"
    ref2 = json.loads(open(os.path.join(root, "profiles", "r02_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref2["impl"] == "reference" and ref2["cpu_baseline"]["kind"] == "reference" and ref2["gpu_launches"] == 0
