"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (run here, no GPU needed):
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAPF = TMA load/store/prefetch, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, UCGABAR = cluster barrier, HMMA = legacy mma.sync (must be absent)."""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quantized_neural_nets_b200", "libgpfq_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR",
        "HMMA.", "DFMA", "FFMA", "LDGSTS"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"\b" + re.escape(k), line):
            cur[k] += 1
print(f"# cuobjdump -sass {os.path.basename(so)} (sm_100a): instruction counts per kernel")
print(f"{'kernel':64s} " + " ".join(f"{k.rstrip('.'):>8s}" for k in KEYS))
for name, c in counts.items():
    if any(c[k] for k in KEYS[:11]) or "gpfq" in name:
        print(f"{name[-64:]:64s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
