"""profiles/r02_k2_traffic.json from an ncu launch list of one K2 pass:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
        --log-file gpurun_out/r02_k2_launches_dram.csv python scripts/prof_k2.py 120 1
    python scripts/ncu_k2_traffic.py gpurun_out/r02_k2_launches_dram.csv 119 > profiles/r02_k2_traffic.json

Sums DRAM bytes and durations over the Farneback launches of the FIRST pass (pyr .. magstats)."""
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = {h: i for i, h in enumerate(rows[start])}
launches = {}
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[hdr["Kernel Name"]].split("(")[0].split("::")[-1]
    L = launches.setdefault(int(r[hdr["ID"]]), {"kernel": name})
    L[r[hdr["Metric Name"]]] = float(r[hdr["Metric Value"]].replace(",", "")) * UNIT[r[hdr["Metric Unit"]]]
order = sorted(launches)
first = next(i for i in order if launches[i]["kernel"].startswith("pyr_kernel"))
last = next(i for i in order if i > first and launches[i]["kernel"].startswith("magstats"))
per = {}
tot = {"read": 0.0, "write": 0.0, "us": 0.0, "launches": 0}
for i in order:
    if not first <= i <= last:
        continue
    L = launches[i]
    k = per.setdefault(L["kernel"], {"read": 0.0, "write": 0.0, "us": 0.0, "launches": 0})
    for d in (k, tot):
        d["read"] += L["dram__bytes_read.sum"]; d["write"] += L["dram__bytes_write.sum"]; d["us"] += L["gpu__time_duration.sum"]; d["launches"] += 1
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold-cache, serialised launches) of scripts/prof_k2.py 120 1",
       "pairs_per_step": int(sys.argv[2]), "dram_bytes_read_per_step": tot["read"], "dram_bytes_write_per_step": tot["write"],
       "dram_bytes_per_step": tot["read"] + tot["write"], "ncu_us_per_step": tot["us"], "launches_per_step": tot["launches"],
       "per_kernel": {k: {"dram_MB_read": v["read"] / 1e6, "dram_MB_write": v["write"] / 1e6, "ncu_us": v["us"], "launches": v["launches"]} for k, v in per.items()}}
print(json.dumps(out, indent=1))
