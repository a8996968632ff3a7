"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

WANT = ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread ", "launch__occupancy_limit",
        "launch__shared_mem_per_block ", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "sm__icc_request_hit_rate", "sass__inst_executed_local", "sm__cycles_elapsed.max ", "smsp__average_warp_latency_per_inst_issued",
        "smsp__pcsamp_warps_issue_stalled", "sm__pipe_tensor", "sm__throughput.avg.pct")
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print(f"# {rep}")
    samples = {}
    for h, u, v in zip(hdr, units, vals):
        hh = h + " "
        if any(w in hh for w in WANT) and "not_issued" not in h:
            print(f"{h} = {v} {u}")
            if h.startswith("smsp__pcsamp_warps_issue_stalled"):
                samples[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v or 0)
    tot = sum(samples.values()) or 1.0
    print("# stall-sample shares: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(samples.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
    print()
