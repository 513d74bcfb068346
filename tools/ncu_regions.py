#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export of k2_synth: executed warp instructions and stall
samples per code region (fast loop / wrap loop / rest), and the top stall reasons.
usage: tools/ncu_regions.py <source.csv> [sample_channels]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
ins = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), r) for r in body]
# region boundaries: the two F2I.F64.TRUNC (window setup) mark the start of the fast and the wrap loop
marks = [i for i, (s, _, _, _) in enumerate(ins) if "F2I.F64.TRUNC" in s]
votes = [i for i, (s, _, _, _) in enumerate(ins) if "VOTE.ANY" in s]
print("F2I marks", marks, "votes", votes, "n", len(ins))
total = sum(e for _, e, _, _ in ins)
samples = sum(s for _, _, s, _ in ins)
def region(a, b, name):
    e = sum(x[1] for x in ins[a:b]); s = sum(x[2] for x in ins[a:b])
    print(f"{name:28s} lines {a:5d}-{b:5d}  exec {e:12d} ({100*e/total:5.1f}%)  samples {s:8d} ({100*s/samples:5.1f}%)")
    return e
if len(marks) >= 2:
    region(0, marks[0], "prologue+chunk init+vote")
    region(marks[0], marks[1], "fast loop (+ gap)")
    region(marks[1], len(ins), "wrap loop + epilogue")
print("total warp instr", total)
if len(sys.argv) > 2:
    sc = float(sys.argv[2])
    print("thread instr per sample-channel", total * 32 / sc)
# stall reasons overall
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = {h: 0 for h in stall_cols}
for _, _, _, r in ins:
    for h in stall_cols:
        v = r[ix[h]]
        tot[h] += int(v) if v else 0
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:9d} {100*v/max(1,sum(tot.values())):5.1f}%")
# opcode mix weighted by executions
mix = {}
for s, e, _, _ in ins:
    parts = s.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + e
for op, e in sorted(mix.items(), key=lambda kv: -kv[1])[:18]:
    print(f"  {op:10s} {e:12d} {100*e/total:5.1f}%")
