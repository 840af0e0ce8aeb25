#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libssr_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / UTCQMMA ...),
TMA (UTMALDG / UBLKCP), tensor-memory loads (LDTM), legacy mma.sync (HMMA) or only CUDA-core FMAs.
Usage: python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_opcodes.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ssr-speech_b200/libssr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
FAMILIES = [("tcgen05 MMA", r"^UTC[A-Z]*MMA"), ("tcgen05 commit/barrier", r"^UTCBAR"), ("TMEM ld/st", r"^(LDTM|STTM)"),
            ("TMEM alloc", r"^UTCATOM|^UTCALLOC|^UTCDEALLOC"), ("TMA tensor", r"^UTMA(LDG|STG|PF|REDG)"), ("bulk copy", r"^UBLKCP"),
            ("mbarrier", r"^SYNCS"), ("mma.sync", r"^(HMMA|IMMA|DMMA)"), ("FFMA", r"^FFMA"), ("LDG", r"^LDG"), ("LDS", r"^LDS"),
            ("cp.async", r"^LDGSTS"), ("SHFL", r"^SHFL"), ("MUFU", r"^MUFU")]
kern = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        counts[kern]["_total"] += 1
        for name, pat in FAMILIES:
            if re.match(pat, op):
                counts[kern][name] += 1
names = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("| kernel | SASS instrs | " + " | ".join(n for n, _ in FAMILIES) + " |")
print("|---|---|" + "---|" * len(FAMILIES))
for (k, c), nm in sorted(zip(counts.items(), names), key=lambda z: z[1]):
    nm = re.sub(r"\(anonymous namespace\)::|ssrb::|void ", "", nm)
    nm = re.sub(r"\((int|bool|unsigned int)\)", "", nm)
    nm = re.sub(r"\(.*", "", nm)
    print(f"| `{nm}` | {c['_total']} | " + " | ".join(str(c[n]) if c[n] else "" for n, _ in FAMILIES) + " |")
