"""SASS mnemonic counts per kernel of the shipped library (`cuobjdump -sass`): the instructions that prove which hardware
paths the kernels use.  Writes profiles/sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "pq3d_b200", "_C", "libpq3d_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
raw = re.findall(r"Function : (\S+)", sass)
dem = subprocess.run(["cu++filt"] + raw, capture_output=True, text=True).stdout.splitlines() if raw else []
for r, d in zip(raw, dem):
    d = d.replace("(int)", "").replace("(bool)", "")
    names[r] = (d[:d.index(">(") + 1] if ">(" in d else d.split("(")[0]).strip()
WANT = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "HMMA", "UCGABAR")
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                per[cur][w] += 1
# UTCHMMA with a tensor-memory A operand = the TS form (P.V of the attention kernel)
ts = collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
    elif cur and re.search(r"UTCHMMA\s+tmem\[", line):
        ts[cur] += 1
out = ["SASS mnemonic counts per kernel (cuobjdump -sass pq3d_b200/_C/libpq3d_b200.so, sm_100a; tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM,",
       "tcgen05.st -> STTM, TMA -> UTMALDG / UTMASTG; HMMA = legacy mma.sync path: none).  UTCHMMA(tmemA) = TS form, A operand in tensor memory.",
       ""]
tot = collections.Counter()
for k, c in per.items():
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]):
        continue
    tot.update(c)
    extra = f" UTCHMMA(tmemA)={ts[k]}" if ts[k] else ""
    out.append(f"{k[:78]:78s} " + " ".join(f"{w}={c[w]}" for w in sorted(c)) + extra)
out += ["", "library total: " + " ".join(f"{w}={tot[w]}" for w in sorted(tot)), f"HMMA (legacy mma.sync) instructions in the library: {sum(c['HMMA'] for c in per.values())}"]
open(os.path.join(root, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:8]))
