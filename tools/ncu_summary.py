"""Summarise ncu outputs (launch list csv / .ncu-rep raw page) into small text files for profiles/."""
import csv
import re
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg, order, lines = {}, [], []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("himgcu::", "")
        us = float(r[vi].replace(",", "")) / 1e3
        lines.append(f"{name:28s} grid={r[gi]:18s} {us:10.1f} us")
        if name not in agg:
            agg[name] = [0.0, 0]
            order.append(name)
        agg[name][0] += us
        agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    out = ["# per-kernel totals (cold-cache, serialised under ncu: compare SHARES)", f"# total {tot:.1f} us"]
    for k in sorted(agg, key=lambda k: -agg[k][0]):
        out.append(f"{k:28s} launches={agg[k][1]:3d} total={agg[k][0]:10.1f} us share={agg[k][0] / tot:6.3f}")
    return "\n".join(out + ["", "# every launch"] + lines)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append(f"== {re.sub(r'[(].*', '', r[hdr.index('Kernel Name')])} grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for w in WANT:
            if w in hdr:
                out.append(f"   {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    return "\n".join(out)


def stalls(path, top=12):
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    out = []
    for s, e in zip(starts, starts[1:]):
        hdr = rows[s + 1]
        si, srci = hdr.index("# Samples"), hdr.index("Source")
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        body = [r for r in rows[s + 2:e] if len(r) > si and r[si].isdigit()]
        tot = sum(int(r[si]) for r in body) or 1
        agg = {}
        for r in body:
            for i in cols:
                if r[i].isdigit():
                    agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
        out.append(f"== {rows[s][1][:60]}  samples={tot}")
        out.append("   " + ", ".join(f"{k}={v / tot:.2f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
        for r in sorted(body, key=lambda r: -int(r[si]))[:top]:
            st = sorted(((hdr[i], int(r[i])) for i in cols if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
            out.append(f"   {int(r[si]) / tot:6.3f} {r[srci].strip()[:60]:60s} {st}")
    return "\n".join(out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    print({"launches": launches, "raw": raw, "stalls": stalls}[mode](path))
