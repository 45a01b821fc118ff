"""Which Blackwell instructions each kernel of libvacnic_b200.so contains (cuobjdump -sass; runs without a GPU):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
HMMA = mma.sync, LDGSTS = cp.async, LDGMC / REDGMC = multimem (NVSwitch multicast) loads / reductions.
  python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vacnic_b200", "libvacnic_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDGSTS", "LDGMC", "REDGMC", "MUFU.EX2",
         "SYNCS", "UCGABAR", "ACQBULK"]
p = subprocess.Popen(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True)
counts = collections.OrderedDict()
cur = None
for line in p.stdout:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for w in WATCH:
        if re.search(r"\b" + re.escape(w), line):
            counts[cur][w] += 1
p.wait()
dem = subprocess.run(["c++filt"], input="\n".join(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel of `vacnic_b200/libvacnic_b200.so` (sm_100a; `tools/sass_mnemonics.py`, cuobjdump -sass)\n")
print("| kernel | " + " | ".join(WATCH) + " |")
print("|---|" + "---:|" * len(WATCH))
for (name, c), d in zip(counts.items(), dem):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("vb::", "")
    print(f"| `{short[:90]}` | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
