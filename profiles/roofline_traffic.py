#!/usr/bin/env python
"""DRAM bytes per launch of the trace kernels from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv`
(run_profile.sh): python profiles/roofline_traffic.py gpurun_out/<tag>_trace_dram.csv > profiles/roofline_traffic.json
bench.py copies the figure of its dominant kernel into roofline.traffic."""
import collections, csv, json, re, sys
MODE = {"0": "traverse", "1": "shade", "2": "shadow_trace", "3": "secondary_trace", "4": "trace_closest"}
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [set(), 0.0])
for row in csv.DictReader(lines):
    m = re.search(r"trace_kernel<\(?(?:dprt::)?[^0-9]*(\d)", row["Kernel Name"])
    if not m or ", 1>" in row["Kernel Name"].split("(")[0]:
        continue                      # skip the instrumented (COUNT = true) variant
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1.0)
    a = agg[MODE[m.group(1)]]
    a[0].add(row["ID"]); a[1] += v
out = {"n_gpus": 1, "source": sys.argv[1], "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras",
       "kernels": {k: {"launches": len(a[0]), "dram_bytes_per_launch": a[1] / max(1, len(a[0]))} for k, a in agg.items()}}
print(json.dumps(out, indent=1))
