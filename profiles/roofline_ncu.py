#!/usr/bin/env python
"""profiles/roofline_traffic.json from one profiling call (run_profile.sh):
    python profiles/roofline_ncu.py gpurun_out/<tag>_trace_dram.csv gpurun_out/<tag>_trace.ncu-rep [gpurun_out/<tag>_part_mlp.ncu-rep] > profiles/roofline_traffic.json
* DRAM bytes per launch of every trace kernel over ALL launches of the bench command (the --metrics dram__bytes pass),
* L2 throughput (lts__t_bytes / duration), issue-slot utilisation and lanes per instruction of the launches the
  `ncu --set full` pass captured (the first launch of each trace kernel; partition / MLP kernels from the second report).
bench.py copies the entry of its dominant kernel into `roofline` (traffic, l2_gbs, issue_active_pct, lanes_per_inst)."""
import collections, csv, io, json, re, subprocess, sys
MODE = {"0": "traverse", "1": "shade", "2": "shadow_trace", "3": "secondary_trace", "4": "trace_closest"}


def num(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else float("nan")


def stage_of(name):
    m = re.search(r"trace_kernel<\(?(?:dprt::)?[^0-9]*(\d)", name)
    if m and ", 1>" not in name.split("(")[0] and "(bool)1" not in name:
        return MODE[m.group(1)]
    if "partition_kernel" in name:
        return "partition_paths" if "PathOps" in name else "partition_queries"
    if "mlp_kernel" in name:
        return "proxy_mlp"
    return None


def dram_per_launch(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [set(), 0.0])
    for row in csv.DictReader(lines):
        st = stage_of(row["Kernel Name"])
        if st is None:
            continue
        v = num(row["Metric Value"]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1.0)
        agg[st][0].add(row["ID"]); agg[st][1] += v
    return {k: {"launches": len(a[0]), "dram_bytes_per_launch": a[1] / max(1, len(a[0]))} for k, a in agg.items()}


def full_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, key):
        i = ix.get(key)
        if i is None:
            return float("nan")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
                 "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(units[i], 1.0)
        return num(r[i]) * scale
    res = {}
    for r in data:
        st = stage_of(r[ix["Kernel Name"]])
        if st is None or st in res:
            continue
        dur = get(r, "gpu__time_duration.sum")
        res[st] = {"capture_duration_us": dur * 1e6, "l2_gbs": get(r, "lts__t_sectors.sum") * 32.0 / dur / 1e9,
                   "dram_gbs": (get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum")) / dur / 1e9,
                   "issue_active_pct": get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                   "lanes_per_inst": get(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct"), "l1_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"),
                   "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "tensor_pipe_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") if st == "proxy_mlp" else None,
                   "registers": get(r, "launch__registers_per_thread")}
    return res


def main():
    kernels = dram_per_launch(sys.argv[1])
    for rep in sys.argv[2:]:
        for st, m in full_metrics(rep).items():
            kernels.setdefault(st, {}).update({k: v for k, v in m.items() if v is not None and v == v})
            kernels[st]["source"] = rep
    print(json.dumps({"n_gpus": 1, "tris_per_chunk": 1000000, "source": sys.argv[1:],
                      "command": "run_profile.sh: ncu --metrics dram__bytes_{read,write}.sum / ncu --set full, --clock-control none, python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras",
                      "kernels": kernels}, indent=1))


main()
