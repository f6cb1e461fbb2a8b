"""`python profiles/make_ncu_traffic.py <one fused step>.ncu-rep [commit]` -> profiles/ncu_traffic.json: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full --clock-control none) of every kernel of the step, averaged
over the launches of that kernel in the report.  bench.py reads `jacobi_bytes_per_launch` for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE[units[col[name]]]


NAMES = (("k_jacobi_pk", "jacobi"), ("k_jacobi_tb", "jacobi_tb"), ("k_fct_x5", "fct_x"), ("k_fct_y5", "fct_y"), ("k_advect5", "advect"),
         ("k_project4", "project"), ("k_rhs", "rhs"), ("k_kappa5", "kappa"))
acc = {}
for r in data:
    kname = r[col["Kernel Name"]]
    for pat, key in NAMES:
        if pat in kname:
            b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            acc.setdefault(key, []).append((b, t))
out = {"source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, fused steps at 8192^2, -ic 3, developed flow), commit {commit}"}
for key, v in acc.items():
    out[f"{key}_bytes_per_launch"] = round(sum(b for b, _ in v) / len(v))
    out[f"{key}_ncu_us_per_launch"] = round(sum(t for _, t in v) / len(v), 1)
    out[f"{key}_launches_in_report"] = len(v)
out["jacobi_sweeps_per_launch"] = 5
out["jacobi_unblocked_algorithmic_bytes_per_launch"] = 12 * 8192 * 8192 * 5
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
