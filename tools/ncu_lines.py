"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
instructions executed and stall samples.   python tools/ncu_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
out = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None:
        continue
    if r[0].strip().isdigit():
        try:
            inst = int(r[hdr.index("Instructions Executed")] or 0)
            samp = int(r[4] or 0)
        except ValueError:
            continue
        out.append((inst, samp, cur_file, int(r[0]), r[1].strip()[:90]))
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
print(f"total inst {tot_i}  samples {tot_s}")
for inst, samp, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * inst / tot_i:5.1f}% inst {100 * samp / tot_s:5.1f}% stall  {f}:{ln:<4d} {src}")
