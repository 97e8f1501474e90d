"""Extracts the judged metrics from an .ncu-rep (`ncu --set full`) into a small JSON/markdown summary.
usage: ncu_summary.py report.ncu-rep out_prefix [label]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__inst_executed_op_branch.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
launches = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    m = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in KEYS:
        if k in d and d[k] != "":
            try:
                m[k] = float(d[k].replace(",", ""))
            except ValueError:
                m[k] = d[k]
            if u.get(k):
                m[k + ".unit"] = u[k]
    launches.append(m)
k0 = launches[0]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
dram = sum(k0.get(f"dram__bytes_{rw}.sum", 0.0) * scale.get(k0.get(f"dram__bytes_{rw}.sum.unit", "byte"), 1.0) for rw in ("read", "write"))
summary = {"label": label, "report": rep.split("/")[-1], "dram_bytes_per_launch": dram, "launches": launches}
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# ncu --set full summary: {label}\n\nreport `{rep.split('/')[-1]}` (not committed; regenerate with the command in profiles/README.md)\n\n")
    for m in launches:
        f.write(f"## {m['kernel']}  grid {m['grid']} block {m['block']}\n\n| metric | value |\n|---|---|\n")
        for k in KEYS:
            if k in m:
                f.write(f"| `{k}` | {m[k]:.6g} {m.get(k + '.unit', '')} |\n")
        f.write(f"| DRAM read+write per launch | {dram / 1e6:.1f} MB |\n\n")
print(json.dumps({k: v for k, v in k0.items() if not k.endswith('.unit')}, indent=1)[:3000])
