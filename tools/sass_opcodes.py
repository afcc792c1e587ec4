"""Per-kernel counts of the SASS mnemonics that identify the Blackwell paths (B200_PROFILING.md): TMA (UTMALDG / UBLKCP),
legacy tensor core (HMMA / IMMA), tcgen05 (UTC*MMA, LDTM, STTM), clusters (UCGABAR), cp.async (LDGSTS), REDUX.
    python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "xbitops_b200" / "libxbitops_b200.so")
OPS = ["UTMALDG", "UBLKCP", "HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UCGABAR", "LDGSTS", "REDUX", "SYNCS", "LDS", "STS", "LOP3", "PRMT", "FFMA", "I2FP", "SHFL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        if cur not in counts:
            counts[cur] = collections.Counter()
            order.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        counts[cur][next((o for o in OPS if op.startswith(o)), op)] += 1      # UCGABAR_ARV / _WAIT, SYNCS.* etc. by prefix
        counts[cur]["_all"] += 1
print(f"cuobjdump -sass {Path(lib).name}   (sm_100a only; counts of static instructions per kernel)")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{o:>7s}" for o in OPS))
tot = collections.Counter()
for k in order:
    c = counts[k]
    tot.update(c)
    print(f"{k[:78]:78s} {c['_all']:6d} " + " ".join(f"{c[o]:7d}" for o in OPS))
print(f"{'TOTAL':78s} {tot['_all']:6d} " + " ".join(f"{tot[o]:7d}" for o in OPS))
