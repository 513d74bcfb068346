// gpusim_kernels.h - launch interface between the C ABI (gpusim_api.cu) and the kernels.
#ifndef GPUSIM_KERNELS_H
#define GPUSIM_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "gpusim_core.h"

namespace gpusim {

// Everything the kernels read for one generate call.  All pointers are device memory.
// "rel" epochs: rows/nch/x0/out are already offset to the first epoch of the range.
struct DeviceJob {
    const DevRow *rows;      // [n_epochs][16], active channels compacted to the front
    const uint8_t *nch;      // [n_epochs] active channel count
    const double *x0;        // [n_epochs][16] code_phase at epoch start
    double *ck_x;            // code phase at sample j*chunk, at ck_index(ck_e0 + e, channel, j, kc)
    uint16_t *ck_w;          // 1023-chip wraps before sample j*chunk, same index
    const double *dc;        // FLOAT hosts: [n_epochs][16] 512*RN(f_carr*delt)
    const double *cph0;      // FLOAT hosts: [n_epochs][16] 512*carr_phase at epoch start
    double *ck_c;            // FLOAT hosts: 512*carr_phase at sample j*chunk, same index
    const int32_t *lut_wide; // [512] AccWide::table_entry(cos, sin)
    const uint64_t *lut_f32; // [512] AccF32x2::table_entry(cos, sin)
    const int16_t *sin16;    // [512] plain tables for the generic kernel
    const int16_t *cos16;    // [512]
    const uint32_t *negw;    // [33][33] inverted C/A chips per PRN
    uint8_t *out;            // n_epochs * epoch_bytes
    unsigned int *work_counter; // zeroed by K1, handed out by K2's warps (32 chunks per unit)
    int32_t n_epochs;
    int32_t ck_e0;           // epoch of rows[0] inside the checkpoint arrays (a sub-range of the job K1 ran for)
    int32_t n_samples;       // samples per epoch
    int32_t chunk;           // samples per thread chunk (multiple of 8; of 32 when ppe == 0)
    int32_t kc;              // chunks per epoch = ceil(n_samples/chunk)
    int32_t fmt;             // 1 / 8 / 16
    int32_t epoch_bytes;
    int32_t max_active;      // max nch over the range (sizes the per-thread state)
    int32_t force_wrap_path; // test hook: always run the wrap-checking loop
    int32_t accum;           // 0 = AccWide (64-bit IMAD), 1 = AccF32x2 (FFMA2)
    int32_t sm_count;        // persistent K2 grid: one 512-thread block per SM
    int32_t ppe;             // aligned layout: code periods per epoch (100), chunk = period / q; 0 = plain
    int32_t q;               // aligned layout: chunks per code period
    int32_t n_units;         // work units (32 chunks each) of the tuned kernel
    int32_t carrier_float;   // 1: FLOAT_CARR_PHASE host (double carrier phase), 0: integer carrier
    int32_t lean;            // 1: integer carrier runs k2_lean (default), 0: the round-1 k2_synth (kept as a cross-check)
    int32_t lin_nb;          // > 0: low chip rate - runs that qualify take synth_lin with this many chip boundaries (2 or 4)
    double lin_rinv;         // synth_lin: job-wide estimate of 1 / (f_code*delt)
    int32_t float_narrow;    // 1: FLOAT hosts always use the 384-thread build (test hook, see K2Geom)
};

// Checkpoints are stored in blocks of 32 consecutive epochs, epoch fastest: the chain kernel's warps are one
// channel over 32 consecutive epochs, so the 32 lanes' values of checkpoint j are one 256-byte run (a scattered
// store costs the LSU one cycle per 32-byte sector: 0.26 of the chain kernel's 0.44 ms went there, ncu r02).
__host__ __device__ inline size_t ck_index(int e, int k, int j, int kc)
{
    return ((((size_t)(e >> 5) * kMaxChan + (size_t)k) * (size_t)kc + (size_t)j) << 5) + (size_t)(e & 31);
}
__host__ __device__ inline size_t ck_elems(int n_epochs, int kc) { return (size_t)((n_epochs + 31) / 32) * 32 * kMaxChan * (size_t)kc; }

enum class ChainAlgo { Jump = 0, Replay = 1 };
enum class SynthKernel { Tuned32 = 0, Tuned16 = 1, Generic = 2 };

// K0: data words of n_frames navigation frames (60 each); data bits of the rows that reference them
cudaError_t launch_navmsg(const NavFrame *frames, int n_frames, uint32_t *dwrd, cudaStream_t stream);
cudaError_t launch_navbits(DevRow *rows, int n_rows, const uint32_t *dwrd, cudaStream_t stream);
// K0: the five subframes of n_eph ephemerides (50 words each); frames that name their subframes by ephemeris index
cudaError_t launch_eph2sbf(const NavEph *eph, int n_eph, const NavIono &iono, uint32_t *sbf, cudaStream_t stream);
cudaError_t launch_navmsg_ref(const NavFrameRef *frames, int n_frames, const uint32_t *sbf, uint32_t *dwrd, cudaStream_t stream);
// K1: exact code-phase checkpoints for every (epoch, active channel, chunk)
cudaError_t launch_chain(const DeviceJob &job, ChainAlgo algo, cudaStream_t stream);
// K2: samples -> bytes
cudaError_t launch_synth(const DeviceJob &job, SynthKernel which, cudaStream_t stream);
// dynamic shared memory the tuned kernel needs for a given max_active (0 if it cannot run)
size_t synth_smem_bytes(int max_active, int threads);
int synth_threads();

} // namespace gpusim
#endif
