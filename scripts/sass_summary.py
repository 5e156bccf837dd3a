#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library (no GPU needed): which kernels of esr_nerf_b200/libesr_b200.so carry
tcgen05 MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM), tcgen05 commits (UTCBAR), bulk copies (UBLKCP), asynchronous
copies (LDGSTS), mbarrier operations (SYNCS) and reductions to global memory (REDG / RED), next to their registers,
spill stack and static shared memory (`cuobjdump -sass` + `cuobjdump --dump-resource-usage`, the mnemonics
/opt/skills/guides/B200_PROFILING.md names).  Writes a markdown table:

    python scripts/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "esr_nerf_b200", "libesr_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "LDGSTS", "SYNCS", "REDG", "RED", "ATOMG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name: str) -> str:
    """esr::k_foo<1, 2>(args...) -> k_foo<1, 2>"""
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "").replace("esr::", "")
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return name[:cut]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    counts, instr, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            instr[cur] += 1
            op = m.group(1)
            if op in MNEMONICS:
                counts[cur][op] += 1
    usage, cur = {}, None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    names = demangle(sorted(instr))
    rows = []
    for k in instr:
        reg, stack, shared, local = usage.get(k, (0, 0, 0, 0))
        rows.append((short(names[k]), instr[k], reg, stack, shared, [counts[k][m] for m in MNEMONICS]))
    rows.sort(key=lambda r: (-r[5][0], -r[1]))
    tot = [sum(r[5][i] for r in rows) for i in range(len(MNEMONICS))]
    print("# SASS summary of `esr_nerf_b200/libesr_b200.so` (sm_100a) — `python scripts/sass_summary.py`\n")
    print(f"{len(rows)} kernels, {sum(r[1] for r in rows)} SASS instructions.  Totals: "
          + ", ".join(f"{m} {t}" for m, t in zip(MNEMONICS, tot)) + ".\n")
    print("`UTCHMMA` = `tcgen05.mma` (kind::f16), `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st` (TMEM), `UTCBAR` = "
          "`tcgen05.commit`, `UBLKCP` = `cp.async.bulk`, `LDGSTS` = `cp.async`, `SYNCS` = mbarrier operations, `REDG` = "
          "`red.global`.  `stack` = bytes of per-thread stack (spills and local arrays), `smem` = static shared memory "
          "(the chain kernels take their 190+ KB dynamically).\n")
    print("| kernel | SASS instr. | regs | stack B | static smem B | " + " | ".join(MNEMONICS) + " |")
    print("|---|---:|---:|---:|---:|" + "---:|" * len(MNEMONICS))
    for name, n, reg, stack, shared, c in rows:
        print(f"| `{name}` | {n} | {reg} | {stack} | {shared} | " + " | ".join(str(x) if x else "" for x in c) + " |")


if __name__ == "__main__":
    sys.exit(main())
