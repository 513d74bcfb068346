"""Consumer side of the output formats: what the reference's players do with gpssim.bin.

The players need SDR hardware and un-vendored libraries, so they are out of scope - but the way they READ the
file is the format contract of the path's output (gpssim.c:2266-2288), and it can be replayed in numpy:

  bladeRF   player/bladeplayer.c:190-194, :243-251   1-bit: byte -> lut[byte][0..7] = +-amp, MSB first, as
                                                     I0 Q0 I1 Q1 I2 Q2 I3 Q3 (SC16_Q11); 16-bit: file as is
  LimeSDR   player/limeplayer.c:306-312              16-bit: short >> 4 -> 12-bit DAC value
            player/limeplayer.c:336-341              8-bit: signed char << 4
            player/limeplayer.c:352-378              1-bit: the same expand LUT as bladeRF
  HackRF    player/hackplayer.c:53-66                8-bit: the file's bytes are the transfer buffer (I, Q, I, Q ...)

Expanded through those conventions, the 1-bit, 8-bit and 16-bit files of one scenario must describe the same
samples.  CPU: the oracle's bytes; GPU: the CUDA path's bytes through the C ABI.
"""
import numpy as np
import pytest

import oracle_lib
from conftest import load_golden


def expand_1bit_like_the_players(raw: np.ndarray, amp: int = 2047) -> np.ndarray:
    """bladeplayer.c:190-194 / limeplayer.c:355-360: lut[i][k] = ((i >> (7-k)) & 1) ? amp : -amp"""
    lut = np.empty((256, 8), dtype=np.int16)
    for i in range(256):
        for k in range(8):
            lut[i, k] = amp if (i >> (7 - k)) & 1 else -amp
    return lut[raw].reshape(-1)          # memcpy(write_buffer_current, lut[read_buffer[i]], 8 shorts)


def check_consumer_views(gen):
    """gen(table) -> uint8 bytes of the path for that table"""
    t16, _, _ = load_golden("static_int_b16")
    t16 = t16.slice(0, 2)
    s16 = gen(t16).view(np.int16)                         # I0 Q0 I1 Q1 ... (pluto / bladeRF 16-bit mode read this as is)
    s8 = gen(t16.with_format(8)).view(np.int8)
    b1 = gen(t16.with_format(1))
    n = t16.samples_per_epoch
    assert s16.size == 2 * 2 * n and s8.size == 2 * 2 * n and b1.size == 2 * (n // 4)

    # LimeSDR 16-bit mode: 12-bit DAC word = short >> 4; 8-bit mode: signed char << 4 gives the same word << 4
    dac12_from16 = s16 >> 4
    dac16_from8 = s8.astype(np.int16) << 4
    assert np.array_equal(dac16_from8 >> 4, dac12_from16)            # the 8-bit file is the 16-bit file's top 12 bits
    assert np.array_equal(dac16_from8, s16 & ~np.int16(15))
    # HackRF: bytes of the 8-bit file go to the DAC unchanged, I first
    assert np.array_equal(s8[0::2], (s16[0::2] >> 4).astype(np.int8)) and np.array_equal(s8[1::2], (s16[1::2] >> 4).astype(np.int8))

    # bladeRF / LimeSDR 1-bit mode: the expanded stream has the sign pattern of the 16-bit samples ("> 0", gpssim.c:2273)
    exp = expand_1bit_like_the_players(b1)
    assert exp.size == s16.size                                      # n % 4 == 0 here: no samples dropped (gpssim.c:2276)
    assert np.array_equal(exp > 0, s16 > 0)
    # and I/Q interleaving survives: bit 7 of byte 0 is I of sample 0, bit 6 is Q of sample 0
    assert bool(b1[0] & 0x80) == bool(s16[0] > 0) and bool(b1[0] & 0x40) == bool(s16[1] > 0)


def test_players_views_of_the_oracle_bytes_agree():
    check_consumer_views(lambda t: oracle_lib.generate(t))


@pytest.mark.gpu
def test_players_views_of_the_cuda_bytes_agree(gpu_required):
    import gps_sdr_sim_b200 as gs

    def gen(t):
        with gs.GpuSim.for_table(t) as sim:
            return sim.generate_epochs(t)
    check_consumer_views(gen)
