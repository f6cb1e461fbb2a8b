"""`python profiles/launch_list_md.py launches.csv > launches.md` -- per-kernel totals of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv`)."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("vof::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
setup = ("k_check_div_by_const", "k_set_init_F", "k_copy_frame", "k_cal_nu_rho")
step_total = sum(v[1] for k, v in agg.items() if not k.startswith(setup))
print("| kernel | launches | total us | mean us | share of step kernels |\n|---|---|---|---|---|")
for k, (n, t) in agg.items():
    share = "" if k.startswith(setup) else f"{t / step_total:.3f}"
    print(f"| {k} | {n} | {t:.1f} | {t / n:.1f} | {share} |")
