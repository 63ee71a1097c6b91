"""Summarise an ncu report: key raw metrics + the source lines with the most warp-stall samples.
usage: python tools/ncu_src.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__icc_request_hit_rate.pct", "smsp__pcsamp_sample_count", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for i, n in enumerate(h):
    if n in want:
        print(f"{n:75s} {u[i]:10s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for i, r in enumerate(rows):
    if r and ("Source" in r or "# Source" in r[0:2] or any("Sampling" in c for c in r)):
        hdr = i
        break
if hdr is None:
    print("no source page")
    sys.exit(0)
H = rows[hdr]
def col(name):
    for i, c in enumerate(H):
        if c.strip() == name:
            return i
    return None
ci_src = col("Source")
ci_smp = col("Warp Stall Sampling (All Samples)") or col("# Samples") or col("Warp Stall Sampling (All Cycles)")
ci_ns = col("Warp Stall Sampling (Not-issued Samples)")
print("columns:", [c for c in H][:12], "...")
data = []
tot = 0
for r in rows[hdr + 1:]:
    if len(r) <= max(ci_src or 0, ci_smp or 0):
        continue
    try:
        s = float(r[ci_smp])
    except (ValueError, TypeError):
        continue
    tot += s
    data.append((s, r[ci_src][:150]))
data.sort(reverse=True)
print(f"total samples {tot:.0f}")
for s, t in data[:top]:
    print(f"{100 * s / max(tot, 1):6.2f}%  {t}")
