#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / profiles/ quote.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md profiles/name.md] [--title "..."]
                                   [--traffic-json profiles/traffic_rk4_bench.json --n-traj 10000000]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) [blocks/SM]"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) [blocks/SM]"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active % (of peak, while SM active)"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64-pipe warp instructions"),
    ("smsp__inst_executed.sum", "warp instructions total"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per instruction issued [cycles]"),
    ("sass__inst_executed_local_loads", "local-memory load instructions (spills + VM overflow stack)"),
    ("sass__inst_executed_local_stores", "local-memory store instructions"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("sm__cycles_active.avg", "SM active cycles"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else rep
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = [f"# {title}", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append(f"## kernel `{d.get('Kernel Name', '?')}`  (ID {d.get('ID', '?')})")
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        for k, label in KEYS:
            if k in d:
                out.append(f"| {label} (`{k}`) | {d[k]} {u[k]} |")
        stalls = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        out.append("")
        out.append("warp stall reasons (warps stalled per issued instruction): " +
                   ", ".join(f"{n} {v:.2f}" for v, n in stalls[:9]))
        out.append("")
    if "--traffic-json" in sys.argv:   # what bench.py reports as roofline.traffic (DRAM bytes of one launch)
        import json
        d, u = dict(zip(hdr, rows[2])), dict(zip(hdr, units))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        n_traj = int(float(sys.argv[sys.argv.index("--n-traj") + 1])) if "--n-traj" in sys.argv else 0
        json.dump({"kernel": d.get("Kernel Name"), "n_traj": n_traj, "dram_bytes_read": rd, "dram_bytes_write": wr, "unit": "bytes per launch",
                   "kernel_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u["gpu__time_duration.sum"], 1.0),
                   "source": f"ncu --set full --clock-control none, {rep} (one launch): dram__bytes_read.sum + dram__bytes_write.sum; "
                             "reads are the per-CTA partial-sum rows (zero-filled, then RED-updated) and the meta blob; the kernel is "
                             "FP64-pipe bound (1.9e13 flop against 1e7 B)"},
                  open(sys.argv[sys.argv.index("--traffic-json") + 1], "w"), indent=1)
    text = "\n".join(out)
    print(text)
    if md:
        open(md, "w").write(text + "\n")


if __name__ == "__main__":
    main()
