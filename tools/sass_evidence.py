#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of the built library (no GPU needed):
UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
LDGSTS = cp.async, STG.E.ENL2.256 = 256-bit global stores, HMMA = legacy mma.sync (must be absent).
usage: python tools/sass_evidence.py > profiles/r01_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gif_b200", "libgifb200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|LDGSTS|STG\.E\.ENL2\.256|HMMA|HGMMA)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        for k in PAT.findall(line):
            cnt[cur][k] += 1
    names = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence (`cuobjdump -sass gif_b200/libgifb200.so`, sm_100a), instruction counts per kernel\n")
    cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "STG.E.ENL2.256", "HMMA"]
    print("| kernel | " + " | ".join(f"`{c}`" for c in cols) + " |\n|---|" + "---:|" * len(cols))
    for mangled, name in sorted(zip(cnt, names), key=lambda t: t[1]):
        c = cnt[mangled]
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        print(f"| `{short[:70]}` | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    print("\nUTCHMMA = `tcgen05.mma.kind::tf32`, LDTM = `tcgen05.ld`, UTMALDG = TMA tensor loads, UTCBAR = `tcgen05.commit`,"
          " UTCATOMSWS = TMEM alloc/dealloc, SYNCS = mbarrier ops, LDGSTS = `cp.async`, no legacy `HMMA` anywhere.")


if __name__ == "__main__":
    main()
