#!/usr/bin/env python
"""Key figures of every kernel in an .ncu-rep (ncu --page raw --csv): duration, DRAM bytes, L2->SM bytes, instructions,
issue utilisation, tensor / alu pipe activity, registers.   python tools/ncu_summary.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm"), ("smsp__inst_executed.sum", "warp_instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "alu_pipe_pct"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("sm__cycles_active.avg", "sm_cycles"), ("launch__registers_per_thread", "registers"),
        ("sm__icc_request_hit_rate.pct", "icache_hit_pct"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"kernel: {r[kcol][:110]}")
        for key, name in WANT:
            for i, h in enumerate(hdr):
                if h == key or h.endswith("." + key):
                    print(f"  {name:20s} {r[i]:>18s} {units[i]}")
                    break


if __name__ == "__main__":
    main()
