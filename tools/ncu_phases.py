"""Stall samples of one kernel of an .ncu-rep, cumulated over consecutive SASS ranges.

    python tools/ncu_phases.py report.ncu-rep [bucket size in instructions]

Prints, per bucket of N consecutive SASS instructions: share of all samples, executed instructions, the
dominant opcodes and the dominant stall reasons -- enough to see which phase of a long unrolled kernel
the time goes to."""
import collections
import csv
import subprocess
import sys


def main(path, bucket=250):
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for k, hdr_i in enumerate(starts):
        end = starts[k + 1] - 1 if k + 1 < len(starts) else len(rows)
        if hdr_i and rows[hdr_i - 1] and rows[hdr_i - 1][0] == "Kernel Name":
            print("== " + rows[hdr_i - 1][1][:100])
        report(rows[hdr_i], rows[hdr_i + 1:end], bucket)


def report(hdr, rows, bucket):
    body = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
    si, ei, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[si] or 0) for r in body)
    print("%d SASS instructions, %d samples" % (len(body), total))
    for b0 in range(0, len(body), bucket):
        chunk = body[b0:b0 + bucket]
        smp = sum(int(r[si] or 0) for r in chunk)
        ex = sum(int(r[ei] or 0) for r in chunk)
        ops = collections.Counter(r[src].split()[1 if r[src].startswith("@") else 0].split(".")[0] for r in chunk if r[src].strip())
        st = collections.Counter()
        for r in chunk:
            for i, h in stall_cols:
                st[h[6:]] += int(r[i] or 0)
        print("[%5d..%5d] %5.1f%% of samples, %9d warp instr | %s | %s" % (
            b0, b0 + len(chunk), 100.0 * smp / max(total, 1), ex, " ".join("%s:%d" % kv for kv in ops.most_common(6)),
            " ".join("%s:%.0f%%" % (k, 100.0 * v / max(smp, 1)) for k, v in st.most_common(5))))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 250)
