#!/usr/bin/env python3
"""Randomised parity sweep: the CUDA path through the C ABI against the oracle on seeded random tables
that leave the envelopes of the reference's scenarios (epoch lengths from 104 samples to 20 MS/s, 1..16
channels, ragged slots, extreme Doppler and code rates, gains in and outside the tuned range, every
output format, both carrier modes, every kernel option, back-to-back device calls, rows that take their data
bits from device-built navigation frames).

usage: python tools/fuzz_parity.py [cases] [seed]      (B200 box; exits 1 on the first mismatch)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gps_sdr_sim_b200 as gs  # noqa: E402
import oracle_lib  # noqa: E402

rng = np.random.default_rng(1)

N_CHOICES = [104, 1000, 4096, 10000, 65536, 100000, 123456, 123457, 200000, 208000, 250000, 260000,
             260008, 400000, 840000, 1000000, 1640000, 2000000]
BUDGET = 1.2e8     # sample-channels per case: keeps the oracle at a fraction of a second


def random_table():
    N = int(rng.choice(N_CHOICES))
    C = int(rng.integers(1, 17))
    fmt = int(rng.choice([gs.SC01, gs.SC08, gs.SC16]))
    if fmt == gs.SC01 and N % 4:
        fmt = gs.SC16
    mode = int(rng.choice([gs.CARRIER_INT, gs.CARRIER_INT, gs.CARRIER_FLOAT]))
    if N < 10000:
        mode = gs.CARRIER_INT      # a double carrier step must stay inside (-1, 1) cycles per sample
    E = int(max(1, min(rng.integers(1, 70), BUDGET // (N * C))))
    t = gs.synthetic_table(E, N, C, fmt, seed=int(rng.integers(1 << 30)), carrier_mode=mode)
    active = t.cols["prn"] > 0
    kind = rng.integers(0, 8)
    note = "plain"
    if kind == 1:      # spacecraft-size and absurd carrier steps, both signs
        t.cols["carr_phasestep"][:, : min(C, 4)] = rng.choice([-566774, 566774, 2**31 - 1, -2**31, 1, -1, 0], size=(E, min(C, 4)))
        if mode == gs.CARRIER_FLOAT:
            f = rng.uniform(-45000.0, 45000.0, size=(E, min(C, 4)))
            t.cols["f_carr"][:, : min(C, 4)] = f
        note = "extreme carrier"
    elif kind == 2:    # ragged: slots dropping in and out, an empty epoch
        drop = rng.random(size=t.cols["prn"].shape) < 0.25
        t.cols["prn"][drop] = 0
        if E > 2:
            t.cols["prn"][int(rng.integers(E)), :] = 0
        note = "ragged"
    elif kind == 3:    # gains at and beyond the tuned kernel's range
        g = rng.choice([0, 1, 255, 256, 4000, 127], size=t.cols["gain"].shape)
        t.cols["gain"][:] = np.where(active, g, 0)
        note = "gain range"
    elif kind == 4:    # code phase right below the wrap / at zero, code-frequency extremes
        t.cols["code_phase"][:] = np.where(active, rng.choice([0.0, 1022.999999999, 1022.5, 511.5, 1e-9], size=active.shape), 0.0)
        t.cols["f_code"][:] = np.where(active, 1.023e6 + rng.uniform(-60.0, 60.0, size=active.shape), 0.0)
        note = "code-phase edges"
    elif kind == 5:    # nav data bit edges: icode 19 and all-ones / alternating bit words
        t.cols["icode"][:] = np.where(active, rng.choice([0, 19, 18, 10], size=active.shape), 0)
        t.cols["nav_bits"][:] = np.where(active, rng.choice([0, 0xFFFFFFFF, 0xAAAAAAAA, 0x55555555], size=active.shape), 0).astype(np.uint32)
        note = "data-bit edges"
    elif kind == 6 and mode == gs.CARRIER_FLOAT:   # carrier phase at the ends of [0,1)
        t.cols["carr_phase_f"][:] = np.where(active, rng.choice([0.0, np.nextafter(1.0, 0.0), 0.5, 1e-300], size=active.shape), 0.0)
        note = "float carrier edges"
    opts = {}
    if rng.random() < 0.5:
        opts = dict([(("force_slow", 1), ("layout", 1), ("accum", 0), ("chain_replay", 1), ("force_generic", 1),
                      ("pipeline", 2), ("pipeline", 1), ("float_geom", 1), ("lowrate", 0), ("lean", 0))[int(rng.integers(10))]])
        if "layout" in opts and rng.random() < 0.5:
            opts["chunk"] = int(rng.choice([64, 96, 128, 1024, 4096]))
    return t, opts, note


def by_reference(t):
    """the same rows taking their data bits from device-built navigation frames.  Returns (table for the
    oracle: 32 data bits per row cut from the oracle's words, table for the GPU: frame / iword / ibit per row, frames)."""
    import ctypes
    n = int(rng.integers(1, 40))
    frames = np.zeros(n, dtype=gs.NAV_FRAME)
    frames["sbf"] = (rng.integers(0, 1 << 24, (n, 5, 10), dtype=np.uint64) << np.uint64(6)).astype(np.uint32)
    frames["first"] = (rng.integers(0, 1 << 24, (n, 10), dtype=np.uint64) << np.uint64(6)).astype(np.uint32)
    frames["tow"] = rng.integers(0, 100800, n)
    frames["tow_first"] = rng.integers(0, 100800, n)
    frames["wn"] = rng.integers(0, 1024, n)
    olib = oracle_lib.lib()
    olib.oracle_nav_frame.restype = None
    olib.oracle_nav_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    words = np.zeros((n, 60), dtype=np.uint32)
    for i, f in enumerate(frames):
        a, b = np.ascontiguousarray(f["sbf"]), np.ascontiguousarray(f["first"])
        olib.oracle_nav_frame(a.ctypes.data, b.ctypes.data, int(f["tow_first"]), int(f["tow"]), int(f["wn"]), words[i].ctypes.data)
    shape = t.cols["prn"].shape
    nav_frame = rng.integers(0, n, shape).astype(np.int32)
    iword = rng.integers(0, 60, shape).astype(np.int32)
    ibit = rng.integers(0, 30, shape).astype(np.int32)
    bits = np.zeros(shape, dtype=np.uint32)
    for e in range(shape[0]):
        for k in range(shape[1]):
            bits[e, k] = gs.pack_nav_bits(words[nav_frame[e, k]].astype(np.uint64), int(iword[e, k]), int(ibit[e, k]))
    t_oracle = gs.EpochTable(t.samples_per_epoch, t.delt, t.data_format, t.carrier_mode, dict(t.cols, nav_bits=bits))
    t_gpu = gs.EpochTable(t.samples_per_epoch, t.delt, t.data_format, t.carrier_mode,
                          dict(t.cols, nav_frame=nav_frame, iword=iword, ibit=ibit), nav_by_reference=True)
    return t_oracle, t_gpu, frames


GUARD = 4096          # poisoned bytes either side of the caller's device output buffer
POISON = 0xA5


def sweep(cases, seed, verbose=True):
    """Run `cases` random cases; returns the description of the first mismatch, or None.
    Memory safety rides along: the library's device buffers sit between poisoned guard bands (GPUSIM_GUARD=1,
    gpusim_debug_guard_violations) and so does the caller's device output buffer; any byte of a band that
    changes counts as a mismatch."""
    global rng
    rng = np.random.default_rng(seed)
    os.environ["GPUSIM_GUARD"] = "1"
    stream = torch.cuda.Stream()
    for case in range(cases):
        t, opts, note = random_table()
        frames = None
        if rng.random() < 0.2:      # rows by reference into device-built navigation frames
            t_oracle, t, frames = by_reference(t)
            note += " +navref"
            want = oracle_lib.generate(t_oracle)
        else:
            want = oracle_lib.generate(t)
        eb = t.epoch_bytes
        with gs.GpuSim.for_table(t) as sim:
            for k, v in opts.items():
                sim.set_option(k, v)
            if frames is not None:
                sim.nav_build(frames)
            how = int(rng.integers(3))
            if how == 0:
                got = sim.generate_epochs(t)
            elif how == 1:
                chunks = []
                sim.generate_epochs_to_sink(t, lambda mv: chunks.append(bytes(mv)))
                got = np.frombuffer(b"".join(chunks), dtype=np.uint8)
            else:   # device-resident, three unsynchronised calls on a side stream over random sub-ranges
                sim.upload_table(t)
                whole = torch.full((max(16, t.n_epochs * eb) + 2 * GUARD,), POISON, dtype=torch.uint8, device="cuda")
                buf = whole[GUARD:whole.numel() - GUARD]
                buf.zero_()
                cuts = sorted(set([0, t.n_epochs] + [int(x) for x in rng.integers(0, t.n_epochs + 1, size=2)]))
                torch.cuda.synchronize()
                ok_align = all((a * eb) % 16 == 0 for a in cuts[:-1])
                if not ok_align:
                    cuts = [0, t.n_epochs]
                for a, b in zip(cuts[:-1], cuts[1:]):
                    sim.generate_device(a, b - a, buf.data_ptr() + a * eb, buf.numel() - a * eb, stream=stream.cuda_stream)
                stream.synchronize()
                got = buf[: t.n_epochs * eb].cpu().numpy()
                stray = int((whole[:GUARD] != POISON).sum()) + int((whole[whole.numel() - GUARD:] != POISON).sum())
                if stray:
                    return f"case {case}: {stray} bytes written outside the caller's output buffer"
            fast = sim.timing().fast_path
            bad_guard = sim.guard_violations()
        if bad_guard != 0:
            return f"case {case}: gpusim_debug_guard_violations = {bad_guard} ({note}, opts={opts})"
        same = got.size == want.size and np.array_equal(got, want)
        line = (f"case {case:4d}: N={t.samples_per_epoch:8d} E={t.n_epochs:3d} C={t.max_active():2d} fmt={t.data_format:2d} "
                f"mode={'float' if t.carrier_mode else 'int  '} {note:20s} opts={opts} via={('host', 'sink', 'device')[how]} "
                f"tuned={fast} {'ok' if same else 'MISMATCH'}")
        if verbose:
            print(line, flush=True)
        if not same:
            idx = np.flatnonzero(got[: want.size] != want[: got.size])
            return line + f"; first differing byte {idx[0] if idx.size else -1} of {want.size} (epoch {idx[0] // eb if idx.size else -1})"
    return None


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time()
    bad = sweep(cases, seed)
    print(f"{cases} cases requested, seed {seed}: {'0 mismatches' if bad is None else bad}, {time.time() - t0:.0f} s")
    sys.exit(0 if bad is None else 1)


if __name__ == "__main__":
    main()
