"""Key raw metrics of every kernel in an .ncu-rep (issue rate, pipes, shared-memory pipe, stalls)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def main(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s grid=%s block=%s" % (r[hdr.index("Kernel Name")][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for w in WANT:
            if w in hdr:
                print("   %s = %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = []
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print("   stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True) if v >= 0.01))


if __name__ == "__main__":
    main(sys.argv[1])
