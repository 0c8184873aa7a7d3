"""profiles/r02_counters.json from an `ncu --page raw --csv` export of one tracking iteration of the bench workload:
per kernel, the warp instructions and DRAM bytes of ONE launch (bench.py's roofline.traffic / roofline.issue read them).
    python tools/ncu_counters.py raw.csv <workload name> > profiles/r02_counters.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    v = r[col[name]].replace(",", "")
    x = float(v)
    u = units[col[name]].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


out = {}
for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
    name = re.split(r"[<(]", name)[0].split("::")[-1]
    out[name] = {"warp_inst": num(r, "smsp__inst_executed.sum"),
                 "dram_bytes": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
                 "duration_us_under_ncu": num(r, "gpu__time_duration.sum") if "gpu__time_duration.sum" in col else None,
                 "lanes_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")}
json.dump({"workload": sys.argv[2], "source": "ncu --set full --clock-control none, one tracking iteration (tools/profile_step.py)",
           "kernels": out}, sys.stdout, indent=1)
