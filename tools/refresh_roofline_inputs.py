#!/usr/bin/env python3
"""Refresh the three files bench.py's roofline block reads (profiles/fp64_flops.json,
dram_traffic.json, smem_pipe.json) from ncu captures of ONE launch of the lean kernel:

    python tools/refresh_roofline_inputs.py full_set.ncu-rep fp64_counts.csv n_states "label"

``full_set.ncu-rep``: ``ncu --set full``; ``fp64_counts.csv``: ``ncu --csv --metrics
smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum`` of the same command.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_page(rep: str) -> dict:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return {name: (rows[1][i], rows[2][i]) for i, name in enumerate(rows[0])}


def main() -> None:
    rep, counts, n_states, label = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    m = raw_page(rep)

    def val(name, scale=1.0):
        unit, v = m[name]
        v = float(v)
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0) * scale

    ops = {}
    for row in csv.reader(l for l in open(counts) if l.startswith('"')):
        for key in ("dfma", "dmul", "dadd"):
            if any(f"op_{key}_pred_on" in c for c in row):
                ops[key] = ops.get(key, 0.0) + float(row[-1].replace(",", ""))
    kernel_s = val("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[m["gpu__time_duration.sum"][0]]
    flops = 2.0 * ops["dfma"] + ops["dmul"] + ops["dadd"]
    json.dump({"source": f"ncu --metrics smsp__sass_thread_inst_executed_op_{{dfma,dmul,dadd}}_pred_on.sum, {label}",
               "dfma": ops["dfma"], "dmul": ops["dmul"], "dadd": ops["dadd"], "states_in_launch": n_states,
               "executed_fp64_flops_per_state": flops / n_states, "kernel_s_under_ncu": kernel_s,
               "tflops_under_ncu": flops / kernel_s / 1e12},
              open(os.path.join(ROOT, "profiles", "fp64_flops.json"), "w"), indent=1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    json.dump({"source": f"ncu --set full, {label}", "dram_bytes_read": rd, "dram_bytes_write": wr,
               "states_in_launch": n_states, "dram_bytes_per_state": (rd + wr) / n_states},
              open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    json.dump({"source": f"ncu --set full, {label}",
               "lsu_data_pipe_wavefronts_pct_of_peak": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
               "shared_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
               "shared_bank_conflict_excess_wavefronts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
               "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "active_lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "registers_per_thread": val("launch__registers_per_thread"),
               "warp_instructions_per_state": val("smsp__inst_executed.sum") / n_states},
              open(os.path.join(ROOT, "profiles", "smem_pipe.json"), "w"), indent=1)
    print("refreshed profiles/fp64_flops.json, dram_traffic.json, smem_pipe.json")


if __name__ == "__main__":
    main()
