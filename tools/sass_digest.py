"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (no GPU needed):

    python tools/sass_digest.py > profiles/r02_sass_digest.txt

UTC*MMA = tcgen05.mma, UTMALDG / UBLKCP = TMA tensor / bulk copies, LDTM / STTM = tcgen05.ld / st,
REDG = the fire-and-forget accumulator promotions, LDGSTS = cp.async, DFMA / DMMA = fp64 pipe, HMMA would be the
legacy mma.sync path (none expected).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "group-attribution-for-diffusion-models_b200", "csrc", "libgadm.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTC", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "REDG", "LDGSTS", "SYNCS", "MUFU",
            "DFMA", "DADD", "DMMA", "HMMA", "HGMMA", "FFMA", "IMAD"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernel = None
    counts = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernel = re.sub(r"\(.*", "", kernel)
            counts[kernel] = collections.Counter()
            continue
        if kernel is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            counts[kernel]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    counts[kernel][p] += 1
    cols = ["UTC", "UTMALDG", "UBLKCP", "LDTM", "STTM", "REDG", "LDGSTS", "SYNCS", "MUFU", "DFMA", "HMMA", "_total"]
    print(f"SASS digest of {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes); UTC = UTC*MMA (tcgen05.mma)")
    print(f"{'kernel':78s} " + " ".join(f"{c:>8s}" for c in cols))
    for k, c in counts.items():
        if c["_total"] < 8:
            continue
        print(f"{k[:78]:78s} " + " ".join(f"{c[x]:8d}" for x in cols))


if __name__ == "__main__":
    sys.exit(main())
