#!/usr/bin/env python
"""Print selected metrics of an ncu report: python profiles/ncu_raw.py <file.ncu-rep> [regex]"""
import csv, io, re, subprocess, sys
DEFAULT = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|launch__registers|"
           r"launch__grid_size|launch__occupancy_limit|sm__throughput.avg.pct|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|smsp__issue_active.avg.pct|"
           r"smsp__inst_executed.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|lts__throughput.avg.pct|l1tex__throughput.avg.pct|"
           r"smsp__warp_issue_stalled_.*_per_warp_active.pct|sm__inst_executed_pipe_(fma|alu|lsu|fp64|xu|tensor).*sum$|local_(ld|st).sum$|"
           r"sm__pipe_tensor.*cycles_active.avg.pct|lts__t_bytes.sum$|l1tex__t_bytes.sum$|smsp__cycles_active.avg$|sm__cycles_elapsed.max$")
def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else DEFAULT)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print("kernels:", [r[kn][:50] for r in data])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"{h:85s} {units[i]:10s} " + "  ".join(r[i] for r in data))
main()
