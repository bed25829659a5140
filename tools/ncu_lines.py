#!/usr/bin/env python
"""Rank CUDA source lines of one kernel in an .ncu-rep by warp-stall samples.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_ID [TOP]      # KERNEL_ID e.g. ::regex:tc_filter:2
Uses `ncu --page source --csv --print-source sass,cuda`: line rows carry the aggregate over their SASS instructions."""
import csv
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-id", kid],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, data, tot = "", None, [], 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall0 = hdr.index("stall_barrier")
            continue
        if hdr is None or len(r) <= si or r[2] != "-":   # line rows have "-" as address
            continue
        try:
            s = int(r[si])
        except ValueError:
            continue
        tot += s
        stalls = {}
        for j in range(stall0, min(len(r), stall0 + 17)):
            try:
                v = int(r[j])
            except ValueError:
                v = 0
            if v:
                stalls[hdr[j].replace("stall_", "")] = v
        data.append((s, fname, r[0], r[1].strip()[:90], r[ie], stalls))
    print(f"total samples {tot}")
    for s, f, l, t, ie, st in sorted(data, reverse=True)[:top]:
        big = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  {f}:{l:<5} inst={ie:>10}  [{big}]  {t}")


if __name__ == "__main__":
    main()
