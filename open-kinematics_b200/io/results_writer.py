"""Result tables: the reference's wide-form sweep file (format version 3) and a columnar sink for
batches.

File layout follows reference ``cli/io/results_writer.py:43-63`` (standard columns), ``:149-189``
(column order: ``step_index, solver_converged, solver_max_residual, solver_nfev``, metric columns,
then ``<point>_{x,y,z}``), ``:233-368`` (Parquet: per-field ``unit`` metadata, JSON provenance under
the ``kinematics_meta`` schema key) and ``:370-460`` (CSV: ``# key: value`` comment header, then a
plain header row).  Single sweeps are written row by row from the facade's objects; batches are
written straight from the instance-major arrays the device filled, one Arrow column per output
column without a Python loop over states.
"""

from __future__ import annotations

import csv
import hashlib
import json
import time
from pathlib import Path

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

FORMAT_VERSION = "3"
METADATA_KEY = b"kinematics_meta"
STANDARD_COLUMNS = ("step_index", "solver_converged", "solver_max_residual", "solver_nfev")

from ..core.metrics.registry import metric_unit  # noqa: E402,F401  (re-exported)


def point_key_name(key) -> str:
    """Public name of a point key: ``wheel_center`` or ``left_wheel_center`` (reference
    primitives/point_ref.py:92-94)."""
    return key.name.lower()


def file_hash(path) -> str:
    try:
        with open(path, "rb") as f:
            return hashlib.file_digest(f, "sha256").hexdigest()
    except OSError:
        return ""


def provenance(geometry_path=None, sweep_path=None, **extra) -> dict:
    meta = {"format_version": FORMAT_VERSION, "timestamp": str(time.time()), **extra}
    if geometry_path is not None:
        meta["geometry_path"] = str(geometry_path)
        meta["geometry_hash"] = file_hash(geometry_path)
    if sweep_path is not None:
        meta["sweep_path"] = str(sweep_path)
        meta["sweep_hash"] = file_hash(sweep_path)
    return meta


class SweepTable:
    """Wide-form table of one solved sweep: columns in file order, one list of values each."""

    def __init__(self, columns: dict, units: dict, metadata: dict):
        self.columns, self.units, self.metadata = columns, units, metadata

    @classmethod
    def from_evaluated(cls, suspension, evaluated, geometry_path=None, sweep_path=None, **extra) -> "SweepTable":
        """``evaluated``: an ``EvaluatedSweep`` (``core/sweep.py::solve_evaluated_sweep``)."""
        from ..core.metrics.main import AxleMetricRows
        n = len(evaluated.states)
        columns = {
            "step_index": list(range(n)),
            "solver_converged": [bool(s.converged) for s in evaluated.solver_stats],
            "solver_max_residual": [float(s.max_residual) for s in evaluated.solver_stats],
            "solver_nfev": [int(s.nfev) for s in evaluated.solver_stats],
        }
        units = {}
        rows = [r.flat_row() if isinstance(r, AxleMetricRows) else r for r in evaluated.metrics.rows]
        for name in (rows[0].keys() if rows else ()):
            columns[name] = [None if r[name] is None else float(r[name]) for r in rows]
            units[name] = metric_unit(name)
        for key in suspension.output_points():
            if any(key not in st.positions for st in evaluated.states):
                continue
            base = point_key_name(key)
            for a, axis in enumerate("xyz"):
                columns[f"{base}_{axis}"] = [float(st.positions[key].data[a]) for st in evaluated.states]
                units[f"{base}_{axis}"] = "mm"
        return cls(columns, units, provenance(geometry_path, sweep_path, **extra))

    def to_arrow(self) -> pa.Table:
        fields, arrays = [], []
        for name, values in self.columns.items():
            if name == "solver_converged":
                typ = pa.bool_()
            elif name in ("step_index", "solver_nfev"):
                typ = pa.int64()
            else:
                typ = pa.float64()
            arrays.append(pa.array(values, type=typ))
            unit = self.units.get(name)
            fields.append(pa.field(name, typ, metadata={b"unit": unit.encode()} if unit else None))
        table = pa.Table.from_arrays(arrays, schema=pa.schema(fields))
        return table.replace_schema_metadata({METADATA_KEY: json.dumps(self.metadata).encode()})

    def write_parquet(self, path) -> None:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        pq.write_table(self.to_arrow(), path)

    def write_csv(self, path) -> None:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        names = list(self.columns)
        with open(path, "w", newline="") as f:
            for key, value in self.metadata.items():
                f.write(f"# {key}: {value}\n")
            f.write(f"# column_units: {json.dumps(self.units, sort_keys=True)}\n#\n")
            writer = csv.writer(f, lineterminator="\n")
            writer.writerow(names)
            for i in range(len(self.columns["step_index"])):
                writer.writerow(["" if self.columns[c][i] is None else self.columns[c][i] for c in names])

    def write(self, path) -> None:
        suffix = Path(path).suffix.lower()
        if suffix == ".parquet":
            self.write_parquet(path)
        elif suffix == ".csv":
            self.write_csv(path)
        else:
            raise ValueError(f"Unsupported output format '{suffix}'. Supported formats: .parquet, .csv")


def run_sweep(suspension, sweep_config, output_path, geometry_path=None, sweep_path=None):
    """Solve, evaluate and write one sweep (reference cli/commands/sweep.py:39-79)."""
    from ..core.sweep import solve_evaluated_sweep
    evaluated = solve_evaluated_sweep(suspension, sweep_config)
    SweepTable.from_evaluated(suspension, evaluated, geometry_path, sweep_path).write(output_path)
    return evaluated


def batch_table(result, instance_offset: int = 0, with_positions: bool = True) -> pa.Table:
    """Long-form Arrow table of a ``BatchSweepResult``: one row per (instance, step), the reference's
    column order with ``instance_index`` in front.  Columns are strided views of the instance-major
    arrays handed to Arrow without a per-state Python loop; states after a failed step are null
    in every numeric column (NaN in the arrays)."""
    n_inst, n_steps = result.nfev.shape
    rows = n_inst * n_steps
    inst = np.repeat(np.arange(instance_offset, instance_offset + n_inst, dtype=np.int64), n_steps)
    step = np.tile(np.arange(n_steps, dtype=np.int64), n_inst)
    failed = np.where(result.failed_step < 0, n_steps, result.failed_step)
    converged = (np.arange(n_steps)[None, :] < failed[:, None]).reshape(rows)
    fields = [pa.field("instance_index", pa.int64()), pa.field("step_index", pa.int64()),
              pa.field("solver_converged", pa.bool_()), pa.field("solver_max_residual", pa.float64()),
              pa.field("solver_nfev", pa.int64())]
    arrays = [pa.array(inst), pa.array(step), pa.array(converged),
              pa.array(result.max_residual.reshape(rows), from_pandas=True),
              pa.array(result.nfev.reshape(rows).astype(np.int64))]

    def add(name: str, values: np.ndarray, unit: str) -> None:
        fields.append(pa.field(name, pa.float64(), metadata={b"unit": unit.encode()}))
        arrays.append(pa.array(np.ascontiguousarray(values), from_pandas=True))   # NaN -> null

    if result.metrics is not None:
        flat = result.metrics.reshape(rows, -1)
        for c, name in enumerate(result.metric_names):
            add(name, flat[:, c], metric_unit(name))
    if with_positions and result.positions is not None:
        flat = result.positions.reshape(rows, -1)
        for p, key in enumerate(result.point_keys):
            for a, axis in enumerate("xyz"):
                add(f"{point_key_name(key)}_{axis}", flat[:, 3 * p + a], "mm")
    table = pa.Table.from_arrays(arrays, schema=pa.schema(fields))
    meta = provenance(n_instances=str(n_inst), n_steps=str(n_steps))
    return table.replace_schema_metadata({METADATA_KEY: json.dumps(meta).encode()})


def write_batch_parquet(path, result, instances_per_row_group: int = 4096, with_positions: bool = True) -> None:
    """Stream a batch to one Parquet file, a row group per block of instances."""
    from dataclasses import replace
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    n_inst = result.nfev.shape[0]
    writer = None
    try:
        for begin in range(0, max(n_inst, 1), instances_per_row_group):
            sl = slice(begin, min(begin + instances_per_row_group, n_inst))
            part = replace(
                result, positions=None if result.positions is None else result.positions[sl],
                status=result.status[sl], failed_step=result.failed_step[sl], nfev=result.nfev[sl],
                max_residual=result.max_residual[sl], metrics=None if result.metrics is None else result.metrics[sl],
                tangents=None, velocities=None, tangent_health=None, diagnostics=None, jumps=None)
            table = batch_table(part, instance_offset=begin, with_positions=with_positions)
            if writer is None:
                writer = pq.ParquetWriter(path, table.schema)
            writer.write_table(table)
    finally:
        if writer is not None:
            writer.close()
