#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the in-tree library (cuobjdump -sass): which kernels carry tcgen05 / TMA / TMEM
instructions, how large they are, registers per thread.  Writes profiles/sass_summary.txt.

    python tools/sass_summary.py [LIB.so] [OUT.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "CREDUX", "FMNMX3", "FMNMX", "FFMA", "HMMA", "VOTE",
         "ATOM", "RED", "LDG", "STG", "LDS", "STS", "BAR", "ELECT"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "radar_multimodal_radiology_b200", "csrc", "libradar_retrieval.so")
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur]["_total"] += 1
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    lines = [f"# SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; sm_100a)",
             "# kernel | instructions | registers | " + " ".join(WATCH), ""]
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("void ", "").replace("radar::", "")
        counts = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        lines.append(f"{short:60s} inst={c['_total']:6d} regs={regs.get(k, '?'):>3}  {counts}")
    tc = [k for k, c in kernels.items() if c["UTCHMMA"]]
    lines += ["", f"kernels issuing tcgen05.mma (UTCHMMA): {len(tc)}; with TMA loads (UTMALDG): "
              f"{sum(1 for c in kernels.values() if c['UTMALDG'])}; reading TMEM (LDTM): {sum(1 for c in kernels.values() if c['LDTM'])}",
              "no wgmma / HMMA (mma.sync) instruction anywhere: " + str(all(c["HMMA"] == 0 for c in kernels.values()))]
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
