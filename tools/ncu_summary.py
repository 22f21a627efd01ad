#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_lattice_tc.txt [--top 25]

Per profiled launch: duration, DRAM bytes, pipe utilisation, occupancy, stall reasons, and the
instructions with the most stall samples (needs -lineinfo + --import-source on for source lines).
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def page(rep: str, name: str) -> str:
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main() -> int:
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none --import-source on)", ""]
    raw = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units = raw[0], raw[1]
    for li, vals in enumerate(raw[2:]):
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append(f"## launch {li}: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:78s} {d[k]:>18s} {u[k]}")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(d[h]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("  warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
        lines.append("")
    src = page(rep, "source")
    # the source page holds one table per launch, each introduced by a "Kernel Name" row
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(src)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    seen = set()
    for b in blocks:
        if len(b["rows"]) < 2:
            continue
        sig = (b["name"], len(b["rows"]), tuple(b["rows"][1][:3]) if len(b["rows"]) > 1 else ())
        if sig in seen:  # the source page repeats a kernel's table once per view
            continue
        seen.add(sig)
        h = b["rows"][0]
        ix = {n: i for i, n in enumerate(h)}
        if "# Samples" not in ix:
            continue
        data = [r for r in b["rows"][1:] if len(r) == len(h)]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        lines.append(f"## hot instructions: {b['name']} ({tot} samples over {len(data)} SASS instructions)")
        ops = {}
        for r in data:
            t = r[ix["Source"]].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
        lines.append("  executed warp-instructions by opcode: " + ", ".join(f"{k}={v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]))
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            n = int(r[ix["# Samples"]])
            why = ", ".join(f"{s[6:]}={int(r[ix[s]])}" for s in st if int(r[ix[s]]) > 0.15 * n and n)
            lines.append(f"  {100.0 * n / tot:5.1f}%  {r[ix['Source']].strip()[:72]:72s} {why}")
        lines.append("")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {out} ({len(lines)} lines)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
