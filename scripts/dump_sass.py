"""Writes one SASS listing per __global__ kernel of libdebigulator_b200.so into profiles/sass/ (instruction
encodings stripped) plus a README with instruction counts. usage: python scripts/dump_sass.py [round tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "debigulator_b200", "libdebigulator_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in txt.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        funcs[cur].append(f"/*{m.group(1)}*/  {m.group(2).strip()} ;")
os.makedirs(OUT, exist_ok=True)
for f in os.listdir(OUT):
    if f.endswith(".sass") and f.startswith(tag + "_"):  # listings of other rounds stay
        os.remove(os.path.join(OUT, f))
rows = []
for mangled, ins in funcs.items():
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip().split("(")[0].replace("dbg::", "").replace("void ", "").replace("<", "_").replace(">", "")
    if not ins:
        continue
    with open(os.path.join(OUT, f"{tag}_{name}.sass"), "w") as fh:
        fh.write(f"// {mangled}\n" + "\n".join(ins) + "\n")
    cnt = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i.split("  ", 1)[1]).split()[0].split(".")[0] for i in ins)
    keys = ["LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "MATCH", "VOTE", "REDUX", "ATOMG", "ATOMS", "RED", "BAR", "WARPSYNC", "PRMT", "BRA", "LDL", "STL"]
    rows.append((name, len(ins), ", ".join(f"{k} {cnt[k]}" for k in keys if cnt[k])))
with open(os.path.join(OUT, "README.md" if tag == "r01" else f"README_{tag}.md"), "w") as fh:
    fh.write(f"# SASS listings (round {tag[1:]}, sm_100a)\n\nOne file per kernel, from `cuobjdump -sass debigulator_b200/libdebigulator_b200.so` via\n"
             "`scripts/dump_sass.py` (instruction encodings stripped; `scripts/sass_ctrl.py` decodes the scheduling\n"
             "control fields when they are needed). `LDGSTS` is the cp.async global->shared staging of compressed input;\n"
             "no tensor-core or TMA mnemonics appear because nothing on this path is a dense contraction.\n\n"
             "| kernel | instructions | selected mnemonics |\n|---|---|---|\n")
    for name, n, sel in sorted(rows, key=lambda r: -r[1]):
        fh.write(f"| `{name}` | {n} | {sel} |\n")
print(len(rows), "kernels")
