#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG / UTMASTG (tiled TMA load / store),
UBLKCP (1-D bulk copy), SYNCS (mbarrier), plus FFMA2 / HFMA2 / MUFU for the CUDA-core kernels.

    python tools/sass_summary.py [path/to/lib.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rust-birdnet-onnx_b200", "lib", "libbirdnet_b200.so")
MN = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "MUFU", "LDS", "STS", "LDGSTS"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
                break
names = list(counts)
try:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
except OSError:
    pass
print(f"# {os.path.relpath(so, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("# " + " ".join(f"{k:>8}" for k in ["instrs"] + MN) + "  kernel")
tot = collections.Counter()
for n in names:
    c = counts[n]
    tot.update(c)
    short = demangle.get(n, n)
    short = short.split(">(")[0] + ">" if ">(" in short else re.sub(r"\(.*", "", short)
    short = re.sub(r"^void ", "", short)
    print("  " + " ".join(f"{c[k]:>8}" for k in ["_total"] + MN) + "  " + short)
print("# total")
print("  " + " ".join(f"{tot[k]:>8}" for k in ["_total"] + MN))
