"""Summarise an ncu report (.ncu-rep) into profiles/<name>.md + per-kernel JSON lines.  Usage:
   python tools/ncu_summary.py gpurun_out/prof_r01a.ncu-rep profiles/r01a_full"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
]
recs = []
for r in rows[2:]:
    rec = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
    for k, nm in KEYS:
        if k in idx:
            rec[nm] = f"{r[idx[k]]} {units[idx[k]]}".strip()
    recs.append(rec)
with open(out + ".md", "w") as f:
    f.write(f"# ncu --set full summary of `{rep}`\n\n")
    cols = ["kernel"] + [nm for _, nm in KEYS]
    f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
    for rec in recs:
        f.write("| " + " | ".join(str(rec.get(c, "")) for c in cols) + " |\n")
with open(out + ".jsonl", "w") as f:
    for rec in recs:
        f.write(json.dumps(rec) + "\n")
print(open(out + ".md").read())
