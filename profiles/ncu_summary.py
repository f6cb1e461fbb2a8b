"""Summarise an .ncu-rep (read here, no GPU): per kernel duration, DRAM bytes/throughput, L1/L2 hit, occupancy,
top stall reasons.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=float("nan")):
    i = col.get(name)
    if i is None:
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = col.get(name)
    return units[i] if i is not None else ""


want_stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
lines = ["| # | kernel | dur us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | L1 hit % | L2 hit % | occ % | regs | IPC | top stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for k, r in enumerate(data):
    name = r[col["Kernel Name"]].split("(")[0][:40]
    dur = g(r, "gpu__time_duration.sum")
    du = unit("gpu__time_duration.sum")
    dur_us = dur / 1e3 if du in ("ns", "nsecond") else (dur if du in ("us", "usecond") else dur * 1e3)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = g(r, "dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1)
    wr = g(r, "dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1)
    pct = g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    l1 = g(r, "l1tex__t_sector_hit_rate.pct")
    l2 = g(r, "lts__t_sector_hit_rate.pct")
    occ = g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    regs = g(r, "launch__registers_per_thread")
    ipc = g(r, "sm__inst_executed.avg.per_cycle_active")
    st = sorted(((g(r, h, 0.0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in want_stalls), reverse=True)[:3]
    stalls = ", ".join(f"{n} {v:.1f}" for v, n in st)
    lines.append(f"| {k} | {name} | {dur_us:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/dur_us/1e3:.0f} | {pct:.1f} | {l1:.1f} | {l2:.1f} | {occ:.0f} | {regs:.0f} | {ipc:.2f} | {stalls} |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
