"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and against
bytes the reference produced (tests/golden).  Bit-exact - this is integer / byte work."""
import hashlib

import numpy as np
import pytest

import oracle_lib
import gps_sdr_sim_b200 as gs
from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

ALL_GOLDEN = golden_names()     # integer-carrier and FLOAT_CARR_PHASE hosts


def digests(buf, table):
    eb = table.epoch_bytes
    return [hashlib.sha256(buf[e * eb:(e + 1) * eb].tobytes()).hexdigest() for e in range(table.n_epochs)]


@pytest.fixture(scope="module", autouse=True)
def _gpu(gpu_required):
    lib = gs.load_library(build_if_missing=False)   # the prebuilt in-tree sm_100a library
    assert lib.gpusim_abi_version() == 2


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_cuda_reproduces_reference_bytes(name):
    table, want, head = load_golden(name)
    with gs.GpuSim.for_table(table) as sim:
        out = sim.generate_epochs(table)
        t = sim.timing()
    assert digests(out, table) == want
    assert np.array_equal(out[:head.size], head)
    assert t.launches >= 2
    # tuned kernel wherever its documented ranges hold
    assert t.fast_path == (1 if table.samples_per_epoch % 8 == 0 else 0)
    assert np.array_equal(out, oracle_lib.generate(table))


@pytest.mark.parametrize("name", ["static_int_b16", "static_int_b8", "static_int_b1", "nmea_int_1msps_b1",
                                  "satellite_int_b16", "static_float_b16", "nmea_float_1msps_b1",
                                  "satellite_float_b16"])
@pytest.mark.parametrize("opts", [{"force_slow": 1}, {"force_generic": 1}, {"chain_replay": 1}, {"accum": 0},
                                  {"layout": 1}, {"layout": 1, "chunk": 128}, {"layout": 1, "chunk": 2048},
                                  {"layout": 1, "chunk": 96, "accum": 0}, {"layout": 1, "force_slow": 1},
                                  {"float_geom": 1}, {"float_geom": 1, "force_slow": 1},
                                  {"lean": 0}, {"lean": 0, "force_slow": 1}, {"lean": 0, "layout": 1, "chunk": 128}])
def test_cuda_variants_agree(name, opts):
    table, want, _ = load_golden(name)
    table, want = table.slice(0, 5), want[:5]
    with gs.GpuSim.for_table(table) as sim:
        for k, v in opts.items():
            sim.set_option(k, v)
        out = sim.generate_epochs(table)
        t = sim.timing()
    assert digests(out, table) == want
    if "force_generic" in opts:
        assert t.fast_path == 0


def test_sink_delivers_the_same_bytes_in_order():
    table, want, _ = load_golden("static_int_b16")
    chunks = []
    with gs.GpuSim.for_table(table) as sim:
        sim.generate_epochs_to_sink(table, lambda mv: chunks.append(bytes(mv)))
    out = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    assert len(chunks) >= 1 and digests(out, table) == want


def test_sink_with_many_sub_batches():
    # 1 MS/s 16-bit epochs are 400 kB: 260 epochs = 104 MB -> at least two 64 MiB staging rounds
    t = gs.synthetic_table(260, 100000, 9, gs.SC16, seed=11)
    got = bytearray()
    with gs.GpuSim.for_table(t) as sim:
        sim.generate_epochs_to_sink(t, lambda mv: got.extend(mv))
        whole = sim.generate_epochs(t)
    assert np.array_equal(np.frombuffer(bytes(got), dtype=np.uint8), whole)
    sel = [0, 1, 129, 259]
    for e in sel:
        assert np.array_equal(whole[e * t.epoch_bytes:(e + 1) * t.epoch_bytes], oracle_lib.generate(t.slice(e, 1)))


def test_device_resident_path_and_epoch_independence():
    import torch
    table, want, _ = load_golden("circle_int_b8")
    eb = table.epoch_bytes
    with gs.GpuSim.for_table(table) as sim:
        sim.upload_table(table)
        buf = torch.zeros(table.n_epochs * eb, dtype=torch.uint8, device="cuda")
        sim.generate_device(0, table.n_epochs, buf.data_ptr(), buf.numel())
        torch.cuda.synchronize()
        whole = buf.cpu().numpy()
        assert digests(whole, table) == want
        # any sub-range of epochs equals the same byte range of the whole run, on the caller's stream
        part = torch.zeros(4 * eb, dtype=torch.uint8, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            sim.generate_device(6, 4, part.data_ptr(), part.numel(), stream=s.cuda_stream)
        s.synchronize()
        assert np.array_equal(part.cpu().numpy(), whole[6 * eb:10 * eb])
        t = sim.timing()
        assert t.synth_ms > 0 and t.chain_ms > 0


@pytest.mark.parametrize("mode,pipeline", [(gs.CARRIER_INT, 1), (gs.CARRIER_INT, 2), (gs.CARRIER_INT, 0), (gs.CARRIER_FLOAT, 1)])
def test_back_to_back_device_calls_overlap_safely(mode, pipeline):
    """Consecutive generate_device calls are pipelined inside the library (the chain kernel of call
    i+1 runs beside the synthesis kernel of call i, on alternating checkpoint sets).  Seven calls with
    different ranges and sizes issued without any synchronisation, then a re-upload and more calls:
    every range must equal the oracle."""
    import torch
    t = gs.synthetic_table(24, 260000, 9, gs.SC08, seed=11, carrier_mode=mode)
    want = oracle_lib.generate(t)
    eb = t.epoch_bytes
    ranges = [(0, 24), (3, 5), (8, 16), (0, 1), (23, 1), (5, 19), (0, 24)]
    with gs.GpuSim.for_table(t) as sim:
        sim.set_option("pipeline", pipeline)      # 1 = overlap when the previous call is still running, 2 = always
        sim.upload_table(t)
        s = torch.cuda.Stream()
        bufs = [torch.zeros(n * eb, dtype=torch.uint8, device="cuda") for _, n in ranges]
        torch.cuda.synchronize()
        for (first, n), buf in zip(ranges, bufs):
            sim.generate_device(first, n, buf.data_ptr(), buf.numel(), stream=s.cuda_stream)
        # a new table while the calls above may still be running: upload must wait for them
        t2 = gs.synthetic_table(24, 260000, 9, gs.SC08, seed=12, carrier_mode=mode)
        sim.upload_table(t2)
        again = torch.zeros(4 * eb, dtype=torch.uint8, device="cuda")
        sim.generate_device(10, 4, again.data_ptr(), again.numel(), stream=s.cuda_stream)
        sim.generate_device(10, 4, again.data_ptr(), again.numel(), stream=s.cuda_stream)
        s.synchronize()
        for (first, n), buf in zip(ranges, bufs):
            assert np.array_equal(buf.cpu().numpy(), want[first * eb:(first + n) * eb]), (first, n)
        assert np.array_equal(again.cpu().numpy(), oracle_lib.generate(t2.slice(10, 4)))
        tm = sim.timing()
        assert tm.synth_ms > 0 and tm.chain_ms > 0


@pytest.mark.parametrize("fmt", [gs.SC16, gs.SC08, gs.SC01])
def test_synthetic_rows_sixteen_channels_extreme_ranges(fmt):
    t = gs.synthetic_table(6, 260000, 16, fmt, seed=3)
    t.cols["gain"][:, :] = np.where(t.prn > 0, 255, 0)              # top of the tuned kernel's range
    t.cols["carr_phasestep"][:, :4] = [[-566774, 566774, 2**31 - 1, -2**31]]   # spacecraft-size and absurd steps
    want = oracle_lib.generate(t)
    with gs.GpuSim.for_table(t) as sim:
        assert np.array_equal(sim.generate_epochs(t), want)
        assert sim.timing().fast_path == 1
    t.cols["gain"][0, 0] = 4000                                        # outside: generic kernel, still exact
    want = oracle_lib.generate(t)
    with gs.GpuSim.for_table(t) as sim:
        assert np.array_equal(sim.generate_epochs(t), want)
        assert sim.timing().fast_path == 0


def test_ragged_tables_inactive_slots_and_empty_batch():
    t = gs.synthetic_table(8, 100000, 12, gs.SC16, seed=21)
    t.cols["prn"][2, 3] = 0          # a slot freed mid-run (gpssim.c:1640)
    t.cols["prn"][5, :] = 0          # an epoch with no satellites at all: all-zero samples
    t.cols["prn"][6, 1:] = 0         # a single channel
    want = oracle_lib.generate(t)
    with gs.GpuSim.for_table(t) as sim:
        got = sim.generate_epochs(t)
        assert np.array_equal(got, want)
        assert not got[5 * t.epoch_bytes:6 * t.epoch_bytes].any()
        empty = t.slice(0, 0)
        assert sim.generate_epochs(empty).size == 0


def test_errors_are_reported_not_swallowed():
    t = gs.synthetic_table(4, 100000, 8, gs.SC16)
    with gs.GpuSim.for_table(t, max_batch_epochs=2) as sim:
        with pytest.raises(gs.GpuSimError) as e:
            sim.generate_epochs(t)
        assert e.value.status == 3
    bad = gs.synthetic_table(2, 100000, 8, gs.SC16)
    bad.cols["prn"][1, 0] = 40
    with gs.GpuSim.for_table(bad) as sim:
        with pytest.raises(gs.GpuSimError) as e:
            sim.generate_epochs(bad)
        assert e.value.status == 1
    bad = gs.synthetic_table(2, 100000, 8, gs.SC16)
    bad.cols["code_phase"][0, 0] = 1023.0
    with gs.GpuSim.for_table(bad) as sim:
        with pytest.raises(gs.GpuSimError):
            sim.generate_epochs(bad)
    import torch
    with gs.GpuSim.for_table(t) as sim:
        sim.upload_table(t)
        buf = torch.zeros(4 * t.epoch_bytes + 64, dtype=torch.uint8, device="cuda")
        with pytest.raises(gs.GpuSimError):
            sim.generate_device(0, 4, buf.data_ptr() + 4, buf.numel() - 4)     # misaligned
        with pytest.raises(gs.GpuSimError):
            sim.generate_device(2, 4, buf.data_ptr(), buf.numel())             # beyond the table


def test_full_size_config2_properties():
    """BASELINE config 2 shape (2999 epochs x 13 channels, 2.6 MS/s) at full size: properties that do
    not need the oracle to finish - the three formats are views of the same samples, batches are
    independent - plus the oracle on a few sampled epochs."""
    import torch
    E, N = 2999, 260000
    t16 = gs.synthetic_table(E, N, 13, gs.SC16)
    batch = 500
    rng = np.random.default_rng(1)
    sampled = sorted(rng.choice(E, 6, replace=False).tolist())
    sims = {f: gs.GpuSim(N, t16.delt, f, gs.CARRIER_INT, batch) for f in (gs.SC16, gs.SC08, gs.SC01)}
    try:
        b16 = torch.empty(batch * 4 * N, dtype=torch.uint8, device="cuda")
        b8 = torch.empty(batch * 2 * N, dtype=torch.uint8, device="cuda")
        b1 = torch.empty(batch * (N // 4), dtype=torch.uint8, device="cuda")
        for first in range(0, E, batch):
            n = min(batch, E - first)
            sub = t16.slice(first, n)
            for f, buf in ((gs.SC16, b16), (gs.SC08, b8), (gs.SC01, b1)):
                sims[f].upload_table(sub.with_format(f))
                sims[f].generate_device(0, n, buf.data_ptr(), buf.numel())
            torch.cuda.synchronize()
            s16 = b16[:n * 4 * N].view(torch.int16)
            s8 = b8[:n * 2 * N].view(torch.int8)
            assert torch.equal(s8, (s16 >> 4).to(torch.int8))                   # gpssim.c:2281
            bits = (s16 > 0).view(-1, 8).to(torch.uint8)                         # gpssim.c:2273
            weights = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device="cuda")
            assert torch.equal((bits * weights).sum(dim=1).to(torch.uint8), b1[:n * (N // 4)])
            assert int(s16.abs().max()) < 2048                                   # 12-bit DAC range, SURVEY 7.4 #7
            for e in [e for e in sampled if first <= e < first + n]:
                got = b16[(e - first) * 4 * N:(e - first + 1) * 4 * N].cpu().numpy()
                assert np.array_equal(got, oracle_lib.generate(t16.slice(e, 1)))
    finally:
        for s in sims.values():
            s.close()


@pytest.mark.parametrize("channels", [1, 13, 14, 16])
@pytest.mark.parametrize("fmt", [gs.SC16, gs.SC01])
def test_float_carrier_kernel_geometries(channels, fmt):
    """FLOAT_CARR_PHASE hosts have two tuned kernels: 512 threads x runs of 32 samples while the
    per-thread state of all channels fits in shared memory (<= 13 active channels), 384 threads x
    runs of 16 beyond (and with float_geom=1).  Both against the oracle, and against each other."""
    t = gs.synthetic_table(3, 260000, channels, fmt, seed=77 + channels, carrier_mode=gs.CARRIER_FLOAT)
    want = oracle_lib.generate(t)
    outs = []
    for geom in (0, 1):
        with gs.GpuSim.for_table(t) as sim:
            sim.set_option("float_geom", geom)
            outs.append(sim.generate_epochs(t))
            assert sim.timing().fast_path == 1
    assert np.array_equal(outs[0], want)
    assert np.array_equal(outs[1], want)


def test_kernels_stay_inside_their_buffers(gpu_required, monkeypatch):
    """Memory-safety evidence without compute-sanitizer: every device buffer of the context (rows, both checkpoint
    sets, work counters, its output buffer) and the caller's own output buffer sit between poisoned guard bands;
    after full-size jobs through every entry point no guard byte may have changed."""
    import torch
    monkeypatch.setenv("GPUSIM_GUARD", "1")
    G = 4096
    for name, mode_opts in (("static_int_b16", {}), ("static_int_b1", {}), ("nmea_int_1msps_b1", {}),
                            ("satellite_int_b16", {"force_slow": 1}), ("static_float_b16", {}),
                            ("satellite_float_b16", {"float_geom": 1}), ("odd_rate_int_b16", {}),
                            ("static_int_b8", {"lean": 0})):
        table, want, _ = load_golden(name)
        eb = table.epoch_bytes
        with gs.GpuSim.for_table(table) as sim:
            for k, v in mode_opts.items():
                sim.set_option(k, v)
            out = sim.generate_epochs(table)                       # library-owned output buffer, staged copies
            assert digests(out, table) == want
            assert sim.guard_violations() == 0, name
            pad = (-(table.n_epochs * eb)) % 16
            whole = torch.full((table.n_epochs * eb + pad + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
            sim.upload_table(table)
            sim.generate_device(0, table.n_epochs, whole.data_ptr() + G, table.n_epochs * eb + pad)
            torch.cuda.synchronize()
            got = whole[G:G + table.n_epochs * eb].cpu().numpy()
            assert digests(got, table) == want
            assert int((whole[:G] != 0xA5).sum()) == 0 and int((whole[G + table.n_epochs * eb:] != 0xA5).sum()) == 0, name
            assert sim.guard_violations() == 0, name


@pytest.mark.parametrize("n,fmt,nch,opts", [
    (2000000, gs.SC16, 11, {}),                       # config 5 shape: 20 MS/s, linear path with 2 boundaries per run
    (2000000, gs.SC08, 13, {"layout": 1, "chunk": 512}),
    (1600000, gs.SC01, 16, {}),
    (1000000, gs.SC16, 12, {}),                       # 10 MS/s: the 4-boundary build
    (840000, gs.SC08, 9, {"layout": 1, "chunk": 1024}),
    (2000000, gs.SC16, 11, {"lowrate": 0}),           # the per-sample loop on the same rows
])
def test_low_chip_rate_linear_path(n, fmt, nch, opts):
    """>= 8 samples per chip: chips of a run from the exact linear model (synth_lin) - bytes of the oracle."""
    t = gs.synthetic_table(3, n, nch, fmt, seed=n // 1000 + fmt)
    with gs.GpuSim.for_table(t) as sim:
        for k, v in opts.items():
            sim.set_option(k, v)
        out = sim.generate_epochs(t)
        assert sim.timing().fast_path == 1
    assert np.array_equal(out, oracle_lib.generate(t))
