#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/rNN_ncu_full_summary.json:
per kernel, per captured launch: duration, DRAM bytes read / written, tensor-pipe activity, FMA-pipe activity,
issue-slot utilisation, registers.  bench.py reads `traffic` and `pipe_tensor_pct` from it.
  python tools/summarize_ncu.py profiles/r02_ncu_full_summary.json gpurun_out/r2k_full_*.ncu-rep"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "ms",
    "dram__bytes_read.sum": "dram_read_GB",
    "dram__bytes_write.sum": "dram_write_GB",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}
UNIT = {"ms": {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3},
        "GB": {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}}


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    summary = {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        ki = head.index("Kernel Name")
        for r in rows[2:]:
            m = re.search(r"(tp_fused_fwd_kernel|tp_fwd2_kernel|tp_bwd2_kernel|tp_fwd_kernel|tp_bwd_kernel|k_gemm3x|k_hidden_fwd|k_hidden_bwd|k_edge_embed_\w+|k_gate_\w+)", r[ki])
            name = m.group(1) if m else r[ki][:60]
            ent = {}
            for i, k in enumerate(head):
                if k in WANT:
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    key = WANT[k]
                    if key == "ms":
                        v *= UNIT["ms"].get(units[i], 1.0)
                    elif key.endswith("_GB"):
                        v *= UNIT["GB"].get(units[i], 1.0)
                    ent[key] = v
            ent["report"] = rep.split("/")[-1]
            summary.setdefault(name, []).append(ent)
    json.dump(summary, open(out_path, "w"), indent=1)
    for k, v in summary.items():
        for e in v:
            print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in e.items()})


if __name__ == "__main__":
    main()
