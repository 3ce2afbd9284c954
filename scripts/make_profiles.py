"""Regenerates the text / json summaries under profiles/ from the raw files a gpurun call left in gpurun_out/:
bench_<tag>.json, launches_<tag>.csv, prof_<tag>.ncu-rep (see profiles/README.md for the commands).
usage: python scripts/make_profiles.py [F] [tag]"""
import csv
import io
import json
import shutil
import subprocess
import sys

F = int(sys.argv[1]) if len(sys.argv) > 1 else 148
TAG = sys.argv[2] if len(sys.argv) > 2 else "r02"
subprocess.run(f"python scripts/ncu_summary.py gpurun_out/prof_{TAG}.ncu-rep > profiles/ncu_summary_{TAG}.txt", shell=True, check=True)
shutil.copy(f"gpurun_out/bench_{TAG}.json", f"profiles/bench_{TAG}.json")
shutil.copy(f"gpurun_out/launches_{TAG}.csv", f"profiles/launches_{TAG}.csv")
rows = list(csv.reader(open(f"gpurun_out/launches_{TAG}.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("spvo::", "").replace("void ", "").split("<")[0]
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 70 : python bench.py --steps 4 --warmup 3",
       "# = the 7 device-resident steps of the primary metric (10 launches each); cold-cache, serialised: compare SHARES with bench.py's live shares",
       f"{'kernel':24s} {'launches':>8s} {'total_us':>10s} {'us/launch':>10s} {'share':>7s}"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:24s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")
open(f"profiles/launches_{TAG}.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{TAG}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
o = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("spvo::", "").replace("void ", "").split("<")[0]
    t = sum(float(r[ix[m]].replace(",", "")) * mult[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    o.setdefault(f"{name}@F{F}", t)
json.dump(o, open("profiles/ncu_traffic.json", "w"), indent=1)
d = json.loads(open(f"gpurun_out/bench_{TAG}.json").read().strip().splitlines()[-1])
tk = sum(v["ms_per_launch"] * v["launches"] for v in d["kernels"].values())
print({k: (round(v["ms_per_launch"], 4), round(v["ms_per_launch"] * v["launches"] / tk, 3)) for k, v in d["kernels"].items()})
for k in ("value", "ms_per_step", "roofline", "decode_roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
    print(k, d[k])
