"""gps_sdr_sim_b200 - B200-native GPS L1 C/A sample synthesis behind gps-sdr-sim's host.

Only the hot path of the reference (gpssim.c:2190-2288, and the navigation data words it reads, gpssim.c:1467-1547) lives here: the CUDA library
(csrc/ -> libgpusim.so, C ABI in include/gpusim.h) and the thin Python host above it.
"""
from .table import (CARRIER_FLOAT, CARRIER_INT, MAX_CHAN, NAV_EPH, NAV_FRAME, NAV_FRAME_REF, NAV_IONO, SC01, SC08, SC16, EpochTable, epoch_bytes,
                    synthetic_table)
from .api import (GpuSim, GpuSimError, Timing, advance_carrier_f64, ca_code, carrier_lut, library_path, load_library,
                  pack_nav_bits)

__all__ = ["GpuSim", "GpuSimError", "Timing", "EpochTable", "epoch_bytes", "synthetic_table", "ca_code",
           "carrier_lut", "pack_nav_bits", "advance_carrier_f64", "load_library", "library_path", "MAX_CHAN", "SC01", "SC08", "SC16",
           "CARRIER_INT", "CARRIER_FLOAT", "NAV_FRAME", "NAV_EPH", "NAV_IONO", "NAV_FRAME_REF"]
