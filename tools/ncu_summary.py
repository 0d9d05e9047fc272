"""Summarise an .ncu-rep (read offline, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "l1tex__t_sector_hit_rate.pct",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = []
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"== kernel: {name}")
        for i, h in enumerate(hdr):
            if h in WANT:
                lines.append(f"{h:72s} {row[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h2 = src[hi]
    isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    stall = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in src[hi + 1:] if len(r) > iex and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in data) or 1
    lines.append(f"-- sampled SASS lines (total samples {tot}, SASS lines {len(data)})")
    agg = {h2[i]: sum(int(r[i] or 0) for r in data) for i in stall}
    lines.append("stall reasons (share of samples): " + ", ".join(
        f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:25]:
        lines.append(f"{100 * int(r[isamp]) / tot:5.1f}%  exec={int(r[iex]):>10d}  {r[isrc].strip()[:100]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
