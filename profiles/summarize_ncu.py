"""Summarise an ncu report into the few numbers DESIGN.md / bench.py cite.
usage: python profiles/summarize_ncu.py gpurun_out/scan_full.ncu-rep profiles/r01_scan_full.md [json_out]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
lines = [f"# ncu summary of `{rep}`", "", "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |",
         "|---|---|" + "---|" * len(data)]
js = {}
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        vals = [r[i] for r in data]
        lines.append(f"| {w} | {units[i]} | " + " | ".join(v[:60] for v in vals) + " |")
        js[w] = {"unit": units[i], "values": vals}
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    def num(name):
        i = hdr.index(name)
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[units[i]]
        return [float(r[i]) * mult for r in data]
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    per = [a + b for a, b in zip(rd, wr)]
    json.dump({"source": rep, "kernel": data[0][hdr.index("Kernel Name")], "launches": len(data),
               "dram_bytes_per_launch": sum(per) / len(per), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "gpu_time_ms": js["gpu__time_duration.sum"]["values"]}, open(sys.argv[3], "w"), indent=1)
print("\n".join(lines))
