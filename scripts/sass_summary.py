#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove what the built library uses (B200_PROFILING.md): tcgen05 MMAs
(UTCHMMA), tensor-memory loads / stores (LDTM / STTM), bulk-TMA copies (UBLKCP), tcgen05.commit barriers (UTCBAR),
mbarrier ops (SYNCS), packed fp32 pairs (FFMA2 / FMUL2 / FADD2), plain FFMA.  Runs where nvcc/cuobjdump is (no GPU):

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "umnn_b200", "libumnn_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("UTCATOMSWS(alloc)", r"\bUTCATOMSWS"), ("SYNCS", r"\bSYNCS"),
            ("FFMA2", r"\bFFMA2"), ("FMUL2", r"\bFMUL2"), ("FADD2", r"\bFADD2"), ("FFMA", r"\bFFMA\b"), ("HMMA(mma.sync)", r"\bHMMA"),
            ("STG", r"\bSTG"), ("LDG", r"\bLDG"), ("instructions", r"^\s+/\*[0-9a-f]{4,}\*/\s")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                kernels[cur][name] += 1
    names = list(kernels)
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        pass
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(LIB, REPO)}: {len(kernels)} kernels, arch {sorted(set(arch))}")
    print("# kernel | " + " | ".join(n for n, _ in PATTERNS))
    for k, c in kernels.items():
        short = demangle.get(k, k)
        short = re.sub(r"umnn::\(anonymous namespace\)::", "", short)
        short = re.sub(r"\(.*", "", short)[:70]
        print(f"{short:72s} " + " ".join(f"{c.get(n, 0):6d}" for n, _ in PATTERNS))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"{'ALL KERNELS':72s} " + " ".join(f"{tot.get(n, 0):6d}" for n, _ in PATTERNS))
    libs = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    print("# linked libraries: " + ", ".join(sorted(set(re.findall(r"(lib[\w+.-]+)\.so", libs)))))


if __name__ == "__main__":
    main()
