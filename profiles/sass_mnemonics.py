"""`python profiles/sass_mnemonics.py > profiles/r2_sass_mnemonics.md` -- static evidence from the built library (no GPU):
per kernel of libvof.so the SASS instruction count and the mnemonics that show how it moves data (cp.async = LDGSTS,
128-bit global accesses, shuffles, shared-memory loads, queue atomics, barriers) and how it divides (MUFU.RCP + FCHK =
nvcc's IEEE division, DFMA = the fp64 sub-normal path)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "taichi_2d_vof_b200", "libvof.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
keys = ["LDGSTS", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS", "SHFL", "VOTE", "ATOMG", "BAR.SYNC", "MUFU.RCP", "FCHK", "DFMA", "FFMA2", "FMUL2", "FADD2", "FFMA"]
stats = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        stats[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        stats[name]["n"] += 1
        for k in keys:
            if op.startswith(k) and not (k == "FFMA" and op.startswith("FFMA2")):      # FFMA2 / FMUL2 / FADD2: Blackwell packed fp32
                stats[name][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel of libvof.so (sm_100a; `python profiles/sass_mnemonics.py`)\n")
print("| kernel | SASS instr. | regs | static smem B | " + " | ".join(keys) + " |")
print("|---|---|---|---|" + "---|" * len(keys))
for (mangled, c), dem in zip(stats.items(), demangled):
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("vof::", "")
    r, s = regs.get(mangled, ("", ""))
    print(f"| `{short}` | {c['n']} | {r} | {s} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
