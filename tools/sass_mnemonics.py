"""SASS mnemonic counts per kernel of the in-tree libdge_b200.so (`cuobjdump -sass`) -> markdown table on stdout.
Evidence that the hot kernels are tcgen05 / TMA code (UTCHMMA, LDTM, UTMALDG, UTCBAR) and that no legacy HMMA exists.
usage: python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deep-gan-encoders_b200", "dge_b200", "libdge_b200.so")
if not os.path.exists(so):
    import glob
    so = (glob.glob(os.path.join(ROOT, "deep-gan-encoders_b200", "**", "libdge_b200.so"), recursive=True) or [so])[0]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "HMMA", "LDG", "STG", "SHFL", "ATOMG/RED"]
counts, cur, it = collections.OrderedDict(), None, iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"\(.*$", "", cur.replace("(anonymous namespace)::", ""))
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        if op in ("ATOMG", "RED", "REDG"):
            op = "ATOMG/RED"
        if op in cols:
            counts[cur][op] += 1
print("# SASS mnemonic counts per kernel of the in-tree libdge_b200.so (`cuobjdump -sass`, sm_100a; tools/sass_mnemonics.py)\n")
print("`UTCHMMA` = tcgen05.mma kind::f16, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, `UTCBAR` = tcgen05.commit -> mbarrier;")
print("`HMMA` would be legacy mma.sync.  HBM-bound kernels show only vector `LDG`/`STG`, shuffles and atomics.\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
tot = collections.Counter()
for k, c in sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], kv[0])):
    print(f"| `{k}` | " + " | ".join(str(c[x]) for x in cols) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot[x]) for x in cols) + " |")
