"""Summarise one kernel of an ``ncu --set full`` report as JSON (run where ``ncu`` is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/render.ncu-rep [--rays 262144] > profiles/rNN_ncu_render_kernel.json
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.max": "sm_cycles",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_memory_active_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_sm",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar_to_sm_read",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "lsu_shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "lsu_shared_bank_conflicts",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard_pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct": "stall_wait_pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct": "stall_short_scoreboard_pct",
    "smsp__warp_issue_stalled_sleeping_per_warp_active.pct": "stall_sleeping_pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct": "stall_membar_pct",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--rays", type=int, default=0)
ap.add_argument("--note", default="")
ap.add_argument("--row", type=int, default=-1, help="which captured launch of the report (default: the last)")
args = ap.parse_args()
raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], check=True, capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], (rows[2:][args.row])
out = {"report": args.report, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
for h, u, v in zip(hdr, units, vals):
    if h in KEYS:
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        if u in SCALE:
            x *= SCALE[u]
            u = "ms" if u in ("ns", "us", "ms", "s") else "bytes"
        out[KEYS[h]] = x
        if u and u not in ("%",):
            out[KEYS[h] + "_unit"] = u
if "dram_read" in out and "dram_write" in out:
    out["dram_bytes_per_launch"] = out["dram_read"] + out["dram_write"]
if args.rays:
    out["rays_per_launch"] = args.rays
    if "duration" in out:
        out["rays_per_s_under_ncu"] = args.rays / (out["duration"] * 1e-3)
if args.note:
    out["note"] = args.note
print(json.dumps(out, indent=1))
