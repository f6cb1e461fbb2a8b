"""Dynamic instruction mix of one kernel from an .ncu-rep source page:
python profiles/ncu_instmix.py rep.ncu-rep kernel_regex [cells]"""
import csv, io, subprocess, sys
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
cells = float(sys.argv[3]) if len(sys.argv) > 3 else 8192.0 * 8192.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
# the page repeats a header block per kernel instance; take the first instance
blocks = raw.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(blk)))
hdr = rows[1]
ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data)
print(rows[0][1][:80])
print("total warp-instructions", tot, " per cell:", round(tot * 32 / cells, 1))
c = Counter()
for r in data:
    op = r[ia].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    c[o.split(".")[0]] += int(r[ie])
print("  ".join(f"{k} {v * 32 / cells:.1f}" for k, v in c.most_common(16)))
