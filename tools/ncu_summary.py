"""Turns the `ncu --set full` reports in gpurun_out/ into the markdown tables of profiles/rNN_ncu_full_summaries.md.
Usage: python tools/ncu_summary.py "name::path.ncu-rep" ... > profiles/r01_ncu_full_summaries.md"""
import csv, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def table(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = ["| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append(f"| {k} | {d[k][0]} {d[k][1]} |")
    to_b = lambda k: float(d[k][0].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d[k][1], 1)
    traffic = to_b("dram__bytes_read.sum") + to_b("dram__bytes_write.sum")
    t_us = float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(d["gpu__time_duration.sum"][1], 1)
    lines.append(f"| DRAM traffic (read+write) | {traffic / 1e6:.1f} MB = {traffic / t_us / 1e6:.2f} TB/s under ncu |")
    return "\n".join(lines), traffic


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        name, path = arg.rsplit("::", 1)
        t, _ = table(path)
        print(f"## {name}\n\n{t}\n")
