"""Parse an `ncu --set full` capture of bench.py into profiles/ncu_traffic.json (dram bytes per launch by kernel family)
and a text summary of the metrics DESIGN.md quotes.  usage: python scripts/ncu_traffic.py <capture.ncu-rep> <config> <num_samples> <out_txt>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, config, S, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__cluster_size"]


def family(name):
    if "k_lik_flat<(bool)1" in name or "k_lik_flat<1" in name:
        return "lik_main_global"
    if "k_lik_flat" in name:
        return "lik_main"
    if "k_adam_local" in name:
        return "adam_local_z"
    return None


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic, lines = {}, []
for r in data:
    name = r[col["Kernel Name"]]
    fam = family(name)
    lines.append(name)
    for m in want:
        if m in col:
            lines.append(f"    {m:90s} {r[col[m]]:>16s} {units[col[m]]}")
    if fam and "dram__bytes_read.sum" in col:
        b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic.setdefault(fam, []).append(b)
out = {"capture": os.path.basename(rep), "config": config, "num_samples": S, "how": "ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch",
       "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json"), "w"), indent=1)
open(out_txt, "w").write(f"# ncu --set full capture {os.path.basename(rep)} ({config}, S={S}); one launch per kernel\n" + "\n".join(lines) + "\n")
print(json.dumps(out))
