"""Exports the metrics the DESIGN/bench numbers are based on from an .ncu-rep into profiles/<name>_ncu_raw.txt
(and prints the DRAM bytes per QP for profiles/k3_traffic.json).

    python scripts/ncu_summary.py gpurun_out/prof_r1_v7.ncu-rep profiles/r1_v7_ipm_srbd_ncu_raw.txt [n_qps]
"""
import csv
import subprocess
import sys

KEYS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    nqp = float(sys.argv[3]) if len(sys.argv) > 3 else 1776.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = []
    got = {}
    for i, name in enumerate(hdr):
        if name in KEYS or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio"):
            lines.append("%s [%s] = %s" % (name, units[i], vals[i]))
            got[name] = vals[i]
    lines.sort()
    open(out, "w").write("\n".join(lines) + "\n")
    rd, wr = float(got["dram__bytes_read.sum"]), float(got["dram__bytes_write.sum"])
    print("dram bytes per QP: %.1f (read %.3f GB + write %.3f GB over %d QPs)" % ((rd + wr) * 1e9 / nqp, rd, wr, nqp))


if __name__ == "__main__":
    main()
