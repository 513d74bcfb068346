#!/usr/bin/env python3
"""Instruction mix of one kernel of libgpusim.so (cuobjdump -sass), optionally an address range.
usage: tools/sass_mix.py <substring of mangled name> [--dump] [--from 0x..] [--to 0x..]"""
import collections
import re
import subprocess
import sys

LIB = __import__("os").environ.get("GPUSIM_LIB", "gps_sdr_sim_b200/libgpusim.so")


def kernel_sass(pattern):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        if pattern in name:
            return name, p
    raise SystemExit("no kernel matches " + pattern)


def main():
    pat = sys.argv[1]
    lo = int(sys.argv[sys.argv.index("--from") + 1], 16) if "--from" in sys.argv else 0
    hi = int(sys.argv[sys.argv.index("--to") + 1], 16) if "--to" in sys.argv else 1 << 30
    name, body = kernel_sass(pat)
    mix = collections.Counter()
    lines = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", body):
        addr = int(m.group(1), 16)
        if not (lo <= addr <= hi):
            continue
        ins = m.group(2).strip()
        lines.append(f"{addr:05x}  {ins}")
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
        mix[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("DADD", "LDS", "IMAD", "STG", "LDG")) and "." in op else "")] += 1
    print(name, "instructions:", sum(mix.values()))
    for k, v in mix.most_common():
        print(f"  {v:5d} {k}")
    if "--dump" in sys.argv:
        print("\n".join(lines))


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:  # piped into head
        pass
