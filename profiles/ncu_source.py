#!/usr/bin/env python
"""Per-instruction view of one kernel in an ncu report: python profiles/ncu_source.py <rep> <kernel-index> [min_share]
prints SASS with executed count, avg active threads and stall samples; regions are easy to eyeball."""
import csv, io, subprocess, sys
rep, kidx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.reader(io.StringIO("\n".join(blocks[kidx]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in rows[1:])
tot_thr = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in rows[1:])
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in rows[1:])
print(f"# total warp-instr {tot_inst}, thread-instr {tot_thr}, avg threads {tot_thr / max(tot_inst, 1):.2f}, samples {tot_samp}")
for n, r in enumerate(rows[1:]):
    ie = int(r[ix["Instructions Executed"]] or 0)
    print(f"{n:4d} {ie:10d} {100.0 * ie / tot_inst:5.2f}% thr={r[ix['Avg. Threads Executed']]:>5s} samp={int(r[ix['# Samples']] or 0):6d}  {r[ix['Source']].strip()}")
