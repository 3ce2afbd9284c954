"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_summary_rNN.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "inst")]
print(f"# {rep}: one row per captured launch (ncu --set full --clock-control none); cold-cache, serialised")
print("kernel".ljust(22) + "".join(n.rjust(16) for _, n in want))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("spvo::", "").replace("void ", "").split("<")[0]
    out = name[:21].ljust(22)
    for m, _ in want:
        if m in ix:
            v = r[ix[m]]
            u = units[ix[m]]
            try:
                out += f"{float(v.replace(',', '')):.4g} {u[:6]}".rjust(16)
            except ValueError:
                out += v[:15].rjust(16)
        else:
            out += "-".rjust(16)
    print(out)
