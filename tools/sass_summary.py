"""SASS evidence per kernel of libfocal_b200.so (runs on the CPU box: cuobjdump only): counts of the mnemonics that show
tcgen05 (UTCHMMA), tensor memory (LDTM / STTM), bulk TMA (UBLKCP), mbarrier traffic (SYNCS / UTCBAR), packed fp32
(FFMA2 / FADD2 / FMUL2), MUFU, and the absence of legacy HMMA.   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "focal_b200", "libfocal_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU", "SHFL", "F2FP", "HMMA",
        "LDG", "STG", "UBLKPF"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                total[k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("# " + " ".join(f"{k:>7s}" for k in ["instrs"] + KEYS) + "  kernel")
for (name, c), d in zip(counts.items(), dem):
    short = re.sub(r"\(fb::Plan.*", "", d).replace("void fb::", "").replace("(int)", "")
    print("  " + " ".join(f"{c[k]:7d}" for k in ["_all"] + KEYS) + "  " + short[:90])
print("# total " + " ".join(f"{k}={total[k]}" for k in KEYS))
